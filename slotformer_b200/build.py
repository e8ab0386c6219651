"""Builds libsfb200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m slotformer_b200.build [--force] [--debug]

--debug additionally builds libsfb200_debug.so (-DSFB_DEBUG): the same kernels plus the tcgen05 self-test,
the timeline buffer hook and the SFB_DBG kernel switches used by scripts/prof_*.py.  The product library has
none of those.

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
LIB_DEBUG = os.path.join(LIBDIR, 'libsfb200_debug.so')
SOURCES = ["capi.cu", "sa_pass.cu", "sa_pass_tc.cu", "enc_tail.cu", "sa_update.cu", "ro_kernel.cu", "ro_umma.cu", "ro_pack.cu", "decode_combine.cu",
           "sa_pass_split.cu", "transition.cu"]
DEBUG_SOURCES = ["umma_test.cu"]
HEADERS = ['common.cuh', 'sa_kernel.h', 'ro_kernel.h', 'ro_attn.cuh', 'umma.cuh', 'decode_kernel.h', 'transition_kernel.h', os.path.join('..', '..', 'include', 'sfb200.h')]
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


def _digest(sources, flags):
    h = hashlib.sha256()
    for name in sources + HEADERS:
        with open(os.path.join(CSRC, name), 'rb') as f:
            h.update(f.read())
    h.update(' '.join(flags).encode())
    return h.hexdigest()


def build_extension(force=False, verbose=False, debug=False):
    os.makedirs(LIBDIR, exist_ok=True)
    lib = LIB_DEBUG if debug else LIB
    sources = SOURCES + (DEBUG_SOURCES if debug else [])
    flags = NVCC_FLAGS + (['-DSFB_DEBUG'] if debug else [])
    tag = '_debug' if debug else ''
    stamp = os.path.join(LIBDIR, f'libsfb200{tag}.sha256')
    digest = _digest(sources, flags)
    if not force and os.path.exists(lib) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return lib
    objs = []
    procs = []
    for src in sources:
        obj = os.path.join(LIBDIR, src.replace('.cu', f'{tag}.o'))
        cmd = [_nvcc()] + flags + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(out)
        if pr.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
    cmd = [_nvcc(), '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError(f'link failed:\n{out.stdout}')
    with open(os.path.join(LIBDIR, f'build{tag}.log'), 'w') as f:
        f.write('\n'.join(log))
    with open(stamp, 'w') as f:
        f.write(digest)
    if verbose:
        print('\n'.join(log))
    return lib


if __name__ == '__main__':
    path = build_extension(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(path)
    if '--debug' in sys.argv:
        print(build_extension(force='--force' in sys.argv, verbose='-v' in sys.argv, debug=True))
