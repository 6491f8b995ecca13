"""Compile the CUDA sources in csrc/ into the in-tree C-ABI library (sm_100a only, no JIT cache).

Each translation unit is compiled to an object file under csrc/_obj/ (in parallel, only when it or a header
changed) and the objects are linked into libjarvis_hybridnet_b200.so next to this file."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SUFFIX = os.environ.get("JHN_LIB_SUFFIX", "")                        # suffix: experiment builds
LIB = os.path.join(HERE, "libjarvis_hybridnet_b200%s.so" % SUFFIX)
OBJ = os.path.join(CSRC, "_obj" + SUFFIX)
SOURCES = ["api.cu", "repro.cu", "conv_f32.cu", "conv_tc.cu", "conv3_tc.cu", "head_tc.cu", "tail.cu", "center.cu",
           "ingest.cu", "head2d.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
ABI_VERSION = 6                                                        # == jhn_abi_version() of the sources in csrc/


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(HERE, "..", "include", "jarvis_hybridnet_b200.h"), os.path.abspath(__file__)]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _fingerprint():
    """sha256 over every source / header the library is built from (mtimes do not survive a snapshot copy)."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted([os.path.join(CSRC, s) for s in _sources()] + _headers()):
        h.update(os.path.basename(f).encode() + b"\0" + open(f, "rb").read())
    h.update((" ".join(NVCC_FLAGS) + os.environ.get("JHN_NVCC_EXTRA", "")).encode())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".hash"):
        return True
    return open(LIB + ".hash").read().strip() != _fingerprint()


def build(force=False, verbose=False):
    """nvcc -> libjarvis_hybridnet_b200.so next to this file. Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("JHN_NVCC_EXTRA", "").split()
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    flags_tag = os.path.join(OBJ, "flags.txt")
    flags_now = " ".join(NVCC_FLAGS + extra)
    if not os.path.exists(flags_tag) or open(flags_tag).read() != flags_now:
        force = True

    def compile_one(s):
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return s, 0, ""
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, _sources()))
    bad = [r for r in results if r[1] != 0]
    if bad:
        for s, _, log in bad:
            sys.stderr.write("---- %s ----\n%s\n" % (s, log))
        raise RuntimeError("nvcc failed building " + ", ".join(s for s, _, _ in bad))
    if verbose:
        for s, _, log in results:
            print("---- %s ----\n%s" % (s, log))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in _sources()]
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking " + LIB)
    open(flags_tag, "w").write(flags_now)
    open(LIB + ".hash", "w").write(_fingerprint())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
