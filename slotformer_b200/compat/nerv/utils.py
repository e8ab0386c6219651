"""``nerv.utils`` subset: object (de)serialisation helpers used by the offline
slot-extraction / rollout drivers (reference extract_slots.py:12,58-76)."""
import json
import os
import pickle


def mkdir_or_exist(path):
    os.makedirs(path, exist_ok=True)


def dump_obj(obj, path):
    ext = os.path.splitext(path)[1]
    if ext == '.json':
        with open(path, 'w') as f:
            json.dump(obj, f)
    else:
        with open(path, 'wb') as f:
            pickle.dump(obj, f)


def load_obj(path):
    ext = os.path.splitext(path)[1]
    if ext == '.json':
        with open(path) as f:
            return json.load(f)
    with open(path, 'rb') as f:
        return pickle.load(f)


def strip_suffix(name):
    return os.path.splitext(name)[0]
