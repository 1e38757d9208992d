"""Builds libsag.so (hand-written sm_100a CUDA + the C ABI of include/sag.h) in-tree with nvcc.

Run as `python -m spatialaudiogen_b200.build` or through __graft_entry__.build().  nvcc cross-compiles
without a GPU; the resulting .so is git-ignored but travels to the GPU box with the snapshot.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libsag.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into spatialaudiogen_b200/libsag.so for sm_100a."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    objs = []
    odir = os.path.join(HERE, 'build')
    os.makedirs(odir, exist_ok=True)
    procs = []
    hdrs = glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    for src in sources():
        obj = os.path.join(odir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in [src] + hdrs):
            continue                                      # this object is newer than its source and every header
        cmd = [nvcc, '-c', src, '-o', obj, '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC'] + ARCH
        if verbose:
            cmd += ['-Xptxas', '-v']
        if os.path.basename(src) == 'conv_umma.cu' and os.environ.get('SAG_CONV_REG128'):      # development: register cap of the tcgen05 kernels
            cmd += ['-DSAG_CONV_REG128']
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out.decode())
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed building libsag.so')
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ARCH
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
