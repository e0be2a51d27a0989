"""Build recipes for the native pieces (run by __graft_entry__.build()).

  libbbgpu.so   bboptpy_b200/csrc/*.cu  -> the C-ABI engine (nvcc, sm_100a only)
  _bboptpy*.so  bboptpy_b200/host/*.cpp + bboptpy_b200/py/*.cpp -> pybind11 module exposing the
                reference's Python surface (py/multivariate_py.cpp) over the C ABI

Everything is built IN-TREE so the artefacts travel to the GPU box with the snapshot.
"""
import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbbgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError("build failed: " + cmd[0])
    return r.stdout


def build_engine(force=False):
    cus = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = cus + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "bbgpu.h")]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for cu in cus:
        obj = os.path.join(objdir, os.path.basename(cu)[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, deps):
            procs.append((cu, subprocess.Popen([NVCC] + NVCC_FLAGS + ["-c", cu, "-o", obj], stdout=subprocess.PIPE,
                                               stderr=subprocess.STDOUT, text=True)))
    for cu, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on " + cu)
    if force or procs or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-lcusolver", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
                                                   "-Xlinker", "--no-undefined"])
    return LIB


def build_host(force=False):
    import pybind11
    srcs = sorted(glob.glob(os.path.join(HERE, "host", "*.cpp")) + glob.glob(os.path.join(HERE, "py", "*.cpp")))
    if not srcs:
        return None
    deps = srcs + glob.glob(os.path.join(HERE, "host", "*.h*")) + [os.path.join(ROOT, "include", "bbgpu.h")]
    ext = os.path.join(HERE, "_bboptpy" + sysconfig.get_config_var("EXT_SUFFIX"))
    if force or _newer(ext, deps) or _newer(ext, [LIB]):
        # the system g++: the image's CXX wrapper (/opt/gcc/bin/g++) resolves -lstdc++ to the static archive,
        # and a private copy of libstdc++ inside a Python extension crashes in iostream/locale code
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else os.environ.get("CXX", "g++")
        _run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-I" + pybind11.get_include(),
              "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "host")]
             + srcs + ["-L" + HERE, "-lbbgpu", "-Wl,-rpath,$ORIGIN", "-o", ext])
        deps_out = subprocess.run(["ldd", ext], stdout=subprocess.PIPE, text=True).stdout
        if "libstdc++.so" not in deps_out:
            raise RuntimeError("the pybind module linked libstdc++ statically; refusing to ship it")
    return ext


def build_all(force=False):
    lib = build_engine(force)
    ext = build_host(force)
    return lib, ext


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
