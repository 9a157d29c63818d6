"""Builds libddk.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libddk.so')
SOURCES = ['ddk_api.cu', 'ddk_graph.cu', 'ddk_conv.cu', 'ddk_conv3.cu', 'ddk_conv_tc.cu', 'ddk_conv_tcr.cu', 'ddk_hidden.cu', 'ddk_heads.cu', 'ddk_update.cu']
NVCC_FLAGS = (['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC']
              + os.environ.get('DDK_NVCC_EXTRA', '').split())     # e.g. -DDDK_CONV_TRACE=1 for tools/conv_trace.sh


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), 'include', 'ddk.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit (in parallel) and link the shared library next to this file."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs, procs = [], []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f'--- {src} ---\n{out}\n')
        elif verbose:
            sys.stderr.write(f'--- {src} ---\n{out}\n')
    if failed:
        raise RuntimeError('nvcc failed')
    subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs + ['-lcudart'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
