"""Builds libgnngls_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m gnngls_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  The library is linked against the static CUDA
runtime so it only needs the driver at run time.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
OUT_DIR = os.path.join(HERE, '_lib')
LIB_PATH = os.path.join(OUT_DIR, 'libgnngls_b200.so')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-I', INCLUDE, '-I', CSRC]
# per-translation-unit extra flags; the search/glue units must never contract a*b+c into an FMA
SOURCES = {
    'common.cu': [],
    'search.cu': ['-fmad=false'] + (['-DGLS_STAMPS'] if os.environ.get('GNNGLS_GLS_STAMPS') else []),   # phase timers (tools/gls_stamps.py)
    'glue.cu': ['-fmad=false'],
    'gat.cu': [],
    'gat_kn.cu': [],
    'gat_kn_tc.cu': ['-DKN_STAMPS'] if os.environ.get('GNNGLS_KN_STAMPS') else [],   # phase timers (tools/kn_stamps.py)
    'dense.cu': [],
    'model.cu': [],
}


def nvcc():
    path = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(path):
        raise RuntimeError('nvcc not found; cannot build libgnngls_b200.so')
    return path


HASH_PATH = LIB_PATH + '.srchash'


def source_hash():
    """Content hash of everything the library is built from (sources, headers, flags): travels with the .so so that a
    copy of the tree with different mtimes is not rebuilt, while an edited source always is."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.h', '.cuh')))
    files.append(os.path.join(INCLUDE, 'gnngls_b200.h'))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, 'rb').read())
    h.update(repr((ARCH, COMMON[:3], sorted(SOURCES.items()))).encode())
    return h.hexdigest()


def up_to_date():
    try:
        return os.path.exists(LIB_PATH) and open(HASH_PATH).read().strip() == source_hash()
    except OSError:
        return False


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and up_to_date():
        return LIB_PATH
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers += [os.path.join(INCLUDE, 'gnngls_b200.h'), os.path.abspath(__file__)]
    jobs, objs = [], []
    for src, extra in SOURCES.items():
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            raise RuntimeError('missing source ' + sp)
        obj = os.path.join(OUT_DIR, src.replace('.cu', '.o'))
        objs.append(obj)
        if force or _stale(obj, [sp] + headers):
            jobs.append([nvcc()] + ARCH + COMMON + extra + ['-c', sp, '-o', obj])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            for warn in ex.map(run, jobs):
                if verbose and warn.strip():
                    print(warn)
    if force or jobs or _stale(LIB_PATH, objs):
        run([nvcc()] + ARCH + ['-shared', '-o', LIB_PATH + '.tmp'] + objs)
        os.replace(LIB_PATH + '.tmp', LIB_PATH)             # atomic: a concurrent loader never sees a half-written file
    with open(HASH_PATH, 'w') as f:
        f.write(source_hash())
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
