"""TEST INFRASTRUCTURE — builds oracle/_ref/kaolin_ref_C.so.

Compiles the reference's own hot-path translation units *in place* from
/root/reference (nothing is copied into this repo):

    kaolin/csrc/render/mesh/rasterization.cpp, rasterization_cuda.cu,
    kaolin/csrc/render/mesh/dibr_soft_mask.cpp, dibr_soft_mask_cuda.cu

plus oracle/ref_shim.cpp (our 10-line pybind registration), with the same
optimisation flags the reference's setup.py uses (setup.py:152-163: -O3,
-DWITH_CUDA) but for sm_100a only.  The reference's own build system is not
run.  The result is the reference CUDA path on B200: the primary parity oracle
for `-m gpu` tests and the "reference CUDA" timing row of bench.py.

Only usable where /root/reference exists (this container); the GPU box uses
the prebuilt .so which travels with the snapshot (oracle/_ref is git-ignored
but not gpurun-ignored).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("KAOLIN_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NAME = "kaolin_ref_C"


def build(force=False, verbose=True):
    so = os.path.join(OUT, NAME + ".so")
    csrc = os.path.join(REF, "kaolin", "csrc")
    if not os.path.isdir(csrc):
        if verbose:
            print(f"[build_ref] {csrc} not present; keeping prebuilt {so}"
                  f" ({'found' if os.path.exists(so) else 'MISSING'})")
        return so if os.path.exists(so) else None
    srcs = [
        os.path.join(csrc, "render/mesh/rasterization.cpp"),
        os.path.join(csrc, "render/mesh/rasterization_cuda.cu"),
        os.path.join(csrc, "render/mesh/dibr_soft_mask.cpp"),
        os.path.join(csrc, "render/mesh/dibr_soft_mask_cuda.cu"),
        os.path.join(HERE, "ref_shim.cpp"),
    ]
    if (not force and os.path.exists(so)
            and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs)):
        return so
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")]
    inc += [f"-I{sysconfig.get_paths()['include']}", f"-I{csrc}"]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = ["-O3", "-DWITH_CUDA", "-DTHRUST_IGNORE_CUB_VERSION_CHECK",
              f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-std=c++17"] + inc
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(OUT, os.path.basename(s) + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmd = ["nvcc", "-c", s, "-o", o, "-gencode",
                   "arch=compute_100a,code=sm_100a", "-lineinfo",
                   "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"] + common
        else:
            cmd = ["g++", "-c", s, "-o", o, "-fPIC"] + common
        if verbose:
            print("[build_ref]", " ".join(cmd[:6]), "...")
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference compile failed: " + " ".join(cmd))
    libdirs = ce.library_paths("cuda")
    link = (["g++", "-shared", "-o", so] + objs
            + [f"-L{d}" for d in libdirs]
            + [f"-Wl,-rpath,{d}" for d in libdirs]
            + ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
               "-ltorch_python", "-lcudart"])
    subprocess.check_call(link)
    for o in objs:
        if o.endswith(".cpp.o"):
            os.remove(o)  # keep the .cu.o for SASS inspection
    if verbose:
        print("[build_ref] built", so)
    return so


def load():
    """Import the prebuilt reference extension (needs torch; CUDA to run it)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build(force="--force" in sys.argv)
