"""In-tree nvcc build of csrc/*.cu -> csrc/libhig_b200.so (sm_100a only, -lineinfo for ncu source pages).

Objects are rebuilt only when their source (or a header) is newer, compiled in parallel, and linked with the
static CUDA runtime so the .so only needs the driver at load time.  No libcuda link: the one driver symbol
(cuTensorMapEncodeTiled) is resolved through cudaGetDriverEntryPoint at first use.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libhig_b200.so")
SOURCES = ["capi.cu", "gemm_tcgen05.cu", "gemm2_tcgen05.cu", "gemm_stream.cu", "gemm_simt.cu", "ln_film.cu", "eff_attn.cu", "attn_apply.cu", "attn_apply_tc.cu", "diffusion_ops.cu",
           "joints.cu", "bwd_ops.cu", "eff_attn_bwd.cu", "eff_attn_bwd_tc.cu", "train_ops.cu", "text_ops.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; hig_b200 needs the CUDA toolkit to build its sm_100a kernels")


def _newest_header():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "hig_b200.h"))
    return max(os.path.getmtime(h) for h in hs if os.path.exists(h))


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = _newest_header()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append((src, cmd))
    def run(job):
        src, cmd = job
        p = subprocess.run(cmd, capture_output=True, text=True)
        return src, p.returncode, p.stdout + p.stderr
    failed = False
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, rc, out in ex.map(run, jobs):
            if verbose or rc != 0:
                sys.stderr.write(f"== {src} (rc={rc})\n{out}\n")
            failed |= rc != 0
    if failed:
        raise RuntimeError("nvcc failed; see stderr")
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                      "-cudart", "static", "-Xcompiler", "-fPIC"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
