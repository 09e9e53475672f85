"""Build libb3d.so (hand-written sm_100a CUDA kernels + C ABI) in-tree with nvcc.

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles for
sm_100a without a GPU, so this also is the "does it build" check run by __graft_entry__.build().
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb3d.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math",
] + os.environ.get("B3D_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DB3D_STG_CS for a tuning build (changes the stamp: full rebuild)


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp():
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _obj_stamp(src):
    """Hash of one translation unit: its source, the headers it includes (transitively, from csrc/ and include/), the flags."""
    seen, todo = set(), [src]
    h = hashlib.sha1(" ".join(NVCC_FLAGS).encode())
    inc_dirs = [CSRC, os.path.join(HERE, "..", "include")]
    while todo:
        f = todo.pop()
        if f in seen or not os.path.exists(f):
            continue
        seen.add(f)
        data = open(f, "rb").read()
        h.update(os.path.basename(f).encode())
        h.update(data)
        for line in data.decode(errors="ignore").split("\n"):
            line = line.strip()
            if line.startswith("#include \""):
                name = line.split("\"")[1]
                for d in inc_dirs:
                    todo.append(os.path.join(d, name))
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile the .cu files under csrc/ whose sources / headers changed and link libb3d.so.  Returns the library path."""
    stamp_file = LIB + ".stamp"
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file):
        if open(stamp_file).read().strip() == stamp:
            return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        ostamp = _obj_stamp(src)
        if not force and not verbose and os.path.exists(obj) and os.path.exists(obj + ".stamp") \
                and open(obj + ".stamp").read().strip() == ostamp:
            continue
        if os.path.exists(obj + ".stamp"):
            os.remove(obj + ".stamp")
        cmd = [nvcc] + NVCC_FLAGS + ["-I", CSRC, "-I", os.path.join(HERE, "..", "include"), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), obj, ostamp))
    failed = False
    for src, p, obj, ostamp in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        else:
            with open(obj + ".stamp", "w") as fh:
                fh.write(ostamp)
            if verbose:
                sys.stderr.write(out)
    if failed:
        raise RuntimeError("libb3d build failed")
    cmd = [nvcc, "-shared", "-o", LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("libb3d link failed:\n" + r.stdout)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
