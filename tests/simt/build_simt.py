"""Builds tests/simt/_build/libapk_simt.so: the device part of astrild_b200/csrc/deposit_sorted.cu, unchanged,
compiled by g++ against tests/simt/simt.h (CPU fibers instead of GPU threads).  Tests only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "astrild_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "libapk_simt.so")
HOST_PART_MARK = "static size_t max_bricks(const apk_plan *P) {"
DYN_SMEM_DECL = "extern __shared__ __align__(16) unsigned char smem_raw[];"


def device_part(path: str) -> str:
    """The kernels of a .cu without its host launchers (<<< >>> is not C++) and with the dynamic shared memory
    declaration pointed at the emulator's buffer.  Nothing else is touched."""
    text = open(path).read()
    cut = text.index(HOST_PART_MARK)
    text = text[:cut] + "\n}  // namespace apk\n"
    assert DYN_SMEM_DECL in text
    return text.replace(DYN_SMEM_DECL, "unsigned char *smem_raw = simt::dyn_smem;")


def compile_kernels(kernel_text: str, out_dir: str, extra_flags=()) -> str:
    """g++ build of deposit_host.cpp around the given device source -> <out_dir>/libapk_simt.so"""
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "deposit_sorted_kernels.inc"), "w") as f:
        f.write(kernel_text)
    so = os.path.join(out_dir, "libapk_simt.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w",
           *extra_flags, "-I", out_dir, "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
           os.path.join(HERE, "deposit_host.cpp"), "-o", so]
    subprocess.check_call(cmd)
    return so


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in ("deposit_sorted.cu", "apk_common.cuh", "deposit_common.cuh")]
    srcs += [os.path.join(HERE, f) for f in ("simt.h", "deposit_host.cpp", "build_simt.py")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return SO
    return compile_kernels(device_part(srcs[0]), OUT_DIR)


if __name__ == "__main__":
    print(build(force=True))
