"""Builds libsfb200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m slotformer_b200.build [--force]

The shared library lands in slotformer_b200/lib/ (git-ignored, but it travels to the GPU box
with the repo snapshot).  There is no JIT and no fallback: importing slotformer_b200.engine
without the built library raises.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libsfb200.so')
SOURCES = ["capi.cu", "sa_pass.cu", "sa_update.cu", "ro_kernel.cu", "ro_umma.cu", "umma_test.cu", "decode_combine.cu", "sa_pass_split.cu"]
HEADERS = ['common.cuh', 'sa_kernel.h', 'ro_kernel.h', 'ro_attn.cuh', 'umma.cuh', 'decode_kernel.h', os.path.join('..', '..', 'include', 'sfb200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
              '-Xcompiler', '-fPIC', '--use_fast_math', '-Xptxas', '-v']
# --use_fast_math would change expf/division semantics in the slot update; keep IEEE there
NVCC_FLAGS.remove('--use_fast_math')
if os.environ.get('SFB_FINE_PROF'):      # per-role clock64 trace of one rollout layer (scripts/prof_ro_fine.py)
    NVCC_FLAGS.append('-DSFB_FINE_PROF')


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found')
    return exe


def _digest():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), 'rb') as f:
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_extension(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, 'libsfb200.sha256')
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace('.cu', '.o'))
        cmd = [_nvcc()] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(out)
        if pr.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError(f'link failed:\n{out.stdout}')
    with open(os.path.join(LIBDIR, 'build.log'), 'w') as f:
        f.write('\n'.join(log))
    with open(stamp, 'w') as f:
        f.write(digest)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    path = build_extension(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(path)
