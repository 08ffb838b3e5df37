#!/usr/bin/env python
"""Build libmcd_b200.so with extra nvcc flags into scratch_ab/<name>/ for A/B runs on the GPU box:

    python tools/build_variant.py profile -DMCD_KERNEL_PROFILE
    MCD_B200_LIB=scratch_ab/profile/libmcd_b200.so python tools/ab_configs.py c3 c4

(scratch_ab/ is git-ignored but travels with gpurun.)  Only the host-buffer / ctypes entry points honour
MCD_B200_LIB; the torch operator library always links the in-tree build.
"""
import concurrent.futures
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ROOT, 'scratch_ab', name)
    os.makedirs(out_dir, exist_ok=True)
    units = [(os.path.join(ge.CSRC, src), os.path.join(out_dir, stem + '.o'), f) for src, stem, f in ge.CUDA_UNITS]

    def compile_one(unit):
        src, obj, unit_flags = unit
        subprocess.run([ge._nvcc(), '-std=c++17', '-O3', '-lineinfo', *ge.NVCC_ARCH, *unit_flags, *flags, '-Xcompiler',
                        '-fPIC', '-c', src, '-o', obj], check=True)

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(units)) as pool:
        list(pool.map(compile_one, units))
    out = os.path.join(out_dir, 'libmcd_b200.so')
    subprocess.run([ge._nvcc(), '-shared', *ge.NVCC_ARCH, '-o', out, *[u[1] for u in units]], check=True)
    print(out)


if __name__ == '__main__':
    main()
