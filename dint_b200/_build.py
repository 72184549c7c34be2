"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  dint_b200/lib/libdint_b200.so   CUDA kernels + C ABI (include/dint_b200.h), nvcc, sm_100a only
  dint_b200/lib/libdint_wl.so     workload clients (CPU C++: the reference's closed-loop clients restated)
  dint_b200/lib/dint_udp_server   the reference's UDP server front-end over the C ABI (recvmmsg / sendmmsg)
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
# DINT_LIB_TAG=<tag> (development only, tools/variants.sh): load / build libdint_b200_<tag>.so, compiled with DINT_NVCC_DEFINES
_TAG = os.environ.get("DINT_LIB_TAG", "")
LIB = os.path.join(LIBDIR, f"libdint_b200_{_TAG}.so" if _TAG else "libdint_b200.so")
WL_LIB = os.path.join(LIBDIR, "libdint_wl.so")
UDP_SERVER = os.path.join(LIBDIR, "dint_udp_server")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(exts):
    out = []
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith(exts):
                out.append(os.path.join(root, f))
    return out


def find_nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else None


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    cu_src = _sources((".cu", ".cuh", ".h"))
    if force or _newer(LIB, cu_src):
        nvcc = find_nvcc()
        if nvcc is None:
            raise RuntimeError("nvcc not found: cannot build libdint_b200.so (there is no CPU fallback)")
        # DINT_NVCC_DEFINES: extra -D flags for A/B builds of a development variant (none exist at the moment)
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("DINT_NVCC_DEFINES", "").split() + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", LIB, os.path.join(CSRC, "engine.cu")]
        subprocess.run(cmd, check=True)
    wl_src = [os.path.join(CSRC, "workloads.cc"), os.path.join(CSRC, "txn_workloads.cc")]
    if os.path.exists(wl_src[0]) and (force or _newer(WL_LIB, wl_src + _sources((".h", ".cuh")))):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", WL_LIB] + wl_src, check=True)
    srv_src = os.path.join(CSRC, "udp_server.cc")
    if os.path.exists(srv_src) and (force or _newer(UDP_SERVER, [srv_src, LIB] + _sources((".h",)))):
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", UDP_SERVER, srv_src, "-L" + LIBDIR, "-ldint_b200",
                        "-Wl,-rpath,$ORIGIN", "-lpthread"], check=True)
    return LIB
