"""Builds tests/simt/_build/libapk_simt.so: the device part of astrild_b200/csrc/deposit_sorted.cu, unchanged,
compiled by g++ against tests/simt/simt.h (CPU fibers instead of GPU threads).  Tests only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "astrild_b200", "csrc")
# APK_SIMT_ASAN=1: the same libraries built with AddressSanitizer into _build_asan (run pytest with
# LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0): an out-of-bounds access of a kernel --
# shared memory, particle arrays, meshes, tables -- then aborts the test instead of corrupting a neighbour silently
ASAN = os.environ.get("APK_SIMT_ASAN", "0") not in ("", "0")
OUT_DIR = os.path.join(HERE, "_build_asan" if ASAN else "_build")
SO = os.path.join(OUT_DIR, "libapk_simt.so")
HOST_PART_MARK = "static size_t max_bricks(const apk_plan *P) {"
DYN_SMEM_DECL = "extern __shared__ __align__(16) unsigned char smem_raw[];"


def device_part(path: str, mark: str = HOST_PART_MARK) -> str:
    """The kernels of a .cu without its host launchers (<<< >>> is not C++) and with the dynamic shared memory
    declaration (if any) pointed at the emulator's buffer.  Nothing else is touched."""
    text = open(path).read()
    cut = text.index(mark)
    text = text[:cut] + "\n}  // namespace apk\n"
    return text.replace(DYN_SMEM_DECL, "unsigned char *smem_raw = simt::dyn_smem;")


HOST_MARKS = ("<<<", "APK_CUDA(", "APK_REQUIRE(", "APK_CUFFT(", "set_error(")


def strip_host_functions(text: str) -> str:
    """Drops every top-level definition inside ``namespace apk { ... }`` that is host code (a kernel launch, a CUDA
    runtime call or an error return) and keeps the kernels, device functions, structs and constants as they are."""
    ns = text.index("namespace apk {") + len("namespace apk {")
    head, body = text[:ns], text[ns:]
    out, chunk, depth, i, n = [], [], 0, 0, len(body)
    closed = False
    while i < n:
        c = body[i]
        two = body[i:i + 2]
        if two == "//":
            j = body.index("\n", i) if "\n" in body[i:] else n
            chunk.append(body[i:j]); i = j
            continue
        if two == "/*":
            j = body.index("*/", i) + 2
            chunk.append(body[i:j]); i = j
            continue
        if c in "\"'":
            j = i + 1
            while body[j] != c:
                j += 2 if body[j] == "\\" else 1
            chunk.append(body[i:j + 1]); i = j + 1
            continue
        if c == "{":
            depth += 1
        elif c == "}":
            if depth == 0:            # the namespace's own closing brace
                closed = True
                break
            depth -= 1
        chunk.append(c)
        i += 1
        if depth == 0 and c in "};" and (c == ";" or body[i:i + 1] != ";"):
            piece = "".join(chunk)
            if not any(m in piece for m in HOST_MARKS):
                out.append(piece)
            chunk = []
    assert closed, "namespace apk is not closed"
    return head + "".join(out) + "\n}  // namespace apk\n"


def misc_device_part() -> str:
    """Device code of deposit_atomic.cu, route.cu, mesh_ops.cu and ingest.cu in one translation unit."""
    parts = []
    for f in ("deposit_atomic.cu", "route.cu", "mesh_ops.cu", "ingest.cu"):
        parts.append(f"// ---- {f} ----\n" + strip_host_functions(open(os.path.join(CSRC, f)).read()))
    return "\n".join(parts)


def bin_device_part(path: str) -> str:
    """bin_power.cu without its host launchers; the two inline-PTX RED helpers become plain adds."""
    text = open(path).read()
    text = text[:text.index("size_t bin_smem_bytes(int nedges) {")] + "\n}  // namespace apk\n"
    assert text.count(DYN_SMEM_DECL) == 1
    text = text.replace(DYN_SMEM_DECL, "unsigned char *smem_raw = simt::dyn_smem;")
    for ptx in ('asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");',
                'asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");'):
        assert text.count(ptx) == 1
        text = text.replace(ptx, "*addr += v;")
    return text


def _gxx(src: str, inc_dir: str, so: str, extra_flags=()) -> str:
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w", "-DAPK_SIMT",
           *(("-fsanitize=address", "-fno-omit-frame-pointer") if ASAN else ()), *extra_flags, "-I", inc_dir, "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
           os.path.join(HERE, src), "-o", so]
    subprocess.check_call(cmd)
    return so


def compile_kernels(kernel_text: str, out_dir: str, extra_flags=()) -> str:
    """g++ build of deposit_host.cpp around the given device source -> <out_dir>/libapk_simt.so"""
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "deposit_sorted_kernels.inc"), "w") as f:
        f.write(kernel_text)
    return _gxx("deposit_host.cpp", out_dir, os.path.join(out_dir, "libapk_simt.so"), extra_flags)


def _fresh(so: str, srcs) -> bool:
    return os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs)


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in ("deposit_sorted.cu", "brick_common.cuh", "apk_common.cuh", "deposit_common.cuh")]
    srcs += [os.path.join(HERE, f) for f in ("simt.h", "deposit_host.cpp", "build_simt.py")]
    if not force and _fresh(SO, srcs):
        return SO
    return compile_kernels(device_part(srcs[0]), OUT_DIR)


def build_rot(rot: int) -> str:
    """The same kernels built with -DAPK_TILE_ROT=<rot> (see deposit_sorted.cu); the library's own setting gives build()."""
    text = open(os.path.join(CSRC, "deposit_sorted.cu")).read()
    default = int(text.split("#define APK_TILE_ROT ")[1].split()[0])
    if rot == default:
        return build()
    out_dir = os.path.join(OUT_DIR, f"rot{rot}")
    so = os.path.join(out_dir, "libapk_simt.so")
    srcs = [os.path.join(CSRC, f) for f in ("deposit_sorted.cu", "brick_common.cuh", "apk_common.cuh", "deposit_common.cuh")]
    srcs += [os.path.join(HERE, f) for f in ("simt.h", "deposit_host.cpp", "build_simt.py")]
    if _fresh(so, srcs):
        return so
    return compile_kernels(device_part(srcs[0]), out_dir, (f"-DAPK_TILE_ROT={rot}",))


def build_bin(force: bool = False) -> str:
    """-> tests/simt/_build/libapk_simt_bin.so (bin_power_kernel + bin_fold_kernel on CPU fibers)"""
    so = os.path.join(OUT_DIR, "libapk_simt_bin.so")
    srcs = [os.path.join(CSRC, f) for f in ("bin_power.cu", "apk_common.cuh")]
    srcs += [os.path.join(HERE, f) for f in ("simt.h", "bin_host.cpp", "build_simt.py")]
    if not force and _fresh(so, srcs):
        return so
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "bin_power_kernels.inc"), "w") as f:
        f.write(bin_device_part(srcs[0]))
    return _gxx("bin_host.cpp", OUT_DIR, so)


def kmu_device_part(path: str) -> str:
    """bin_kmu.cu without its C entry points; the two inline-PTX RED helpers become plain adds."""
    text = open(path).read()
    text = text[:text.index("using namespace apk;")]
    for ptx in ('asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");',
                'asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");'):
        assert text.count(ptx) == 1
        text = text.replace(ptx, "*addr += v;")
    return text


def build_kmu(force: bool = False) -> str:
    """-> tests/simt/_build/libapk_simt_kmu.so (bin_kmu_kernel + kmu_fold_kernel on CPU fibers)"""
    so = os.path.join(OUT_DIR, "libapk_simt_kmu.so")
    srcs = [os.path.join(CSRC, f) for f in ("bin_kmu.cu", "apk_common.cuh")]
    srcs += [os.path.join(HERE, f) for f in ("simt.h", "kmu_host.cpp", "build_simt.py")]
    if not force and _fresh(so, srcs):
        return so
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "bin_kmu_kernels.inc"), "w") as f:
        f.write(kmu_device_part(srcs[0]))
    return _gxx("kmu_host.cpp", OUT_DIR, so)


def build_misc(force: bool = False) -> str:
    """-> tests/simt/_build/libapk_simt_misc.so (direct-atomic deposit, slab routing / transpose / ghost adds,
    gridded-field helpers on CPU fibers)"""
    so = os.path.join(OUT_DIR, "libapk_simt_misc.so")
    srcs = [os.path.join(CSRC, f) for f in ("deposit_atomic.cu", "route.cu", "mesh_ops.cu", "ingest.cu", "apk_common.cuh", "deposit_common.cuh")]
    srcs += [os.path.join(HERE, f) for f in ("simt.h", "misc_host.cpp", "build_simt.py")]
    if not force and _fresh(so, srcs):
        return so
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "misc_kernels.inc"), "w") as f:
        f.write(misc_device_part())
    return _gxx("misc_host.cpp", OUT_DIR, so)


if __name__ == "__main__":
    print(build(force=True))
    print(build_bin(force=True))
    print(build_misc(force=True))
    print(build_kmu(force=True))
