"""Build the reference's own native extensions into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference ships two JIT torch extensions on the hot path:
  * ``cd``  -- utils/metrics/distance/cd/chamfer_distance.{cpp,cu}   (CPU ``nnsearch`` + CUDA kernel)
  * ``fps`` -- utils/sampling/fps/furthest_point_sampling.{cpp,cu}   (CUDA only)
This script compiles them *from the sources where they lie under /root/reference* (nothing is
copied into this repository) and leaves only the resulting ``.so`` files under ``oracle/_ref/``.
``oracle/_ref/`` is git-ignored but travels to the GPU box, where the compiled reference kernels
are the strongest parity oracle (reference CUDA code running on the same B200).

Three artefacts:
  dustyref_cd      as shipped: ``load()`` is called with no flags in the reference
                   (cd/chamfer_distance.py:7-13) => g++ -O0 for the CPU twin.
  dustyref_cd_o3   same sources, ``-O3`` (baseline ISA so the .so runs on any host; "honest CPU" arm of BASELINE.md section 3.2).
  dustyref_fps     the PointNet++ FPS/gather kernels (fps/furthest_point_sampling.py:10-16).

Run:  python oracle/build_ref.py            (no-op when /root/reference is absent)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DUSTY_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

TARGETS = {
    "dustyref_cd": dict(
        sources=["utils/metrics/distance/cd/chamfer_distance.cpp",
                 "utils/metrics/distance/cd/chamfer_distance.cu"],
        extra_cflags=[]),
    "dustyref_cd_o3": dict(
        sources=["utils/metrics/distance/cd/chamfer_distance.cpp",
                 "utils/metrics/distance/cd/chamfer_distance.cu"],
        extra_cflags=["-O3"]),
    "dustyref_fps": dict(
        sources=["utils/sampling/fps/furthest_point_sampling.cpp",
                 "utils/sampling/fps/furthest_point_sampling.cu"],
        extra_cflags=[]),
}


def built(name):
    return os.path.exists(os.path.join(OUT, name, name + ".so"))


def build(names=None, verbose=False):
    if not os.path.isdir(REF):
        print(f"[oracle/build_ref] {REF} not present; using prebuilt oracle/_ref if any")
        return False
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    for name, spec in TARGETS.items():
        if names and name not in names:
            continue
        if built(name):
            continue
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name,
             sources=[os.path.join(REF, s) for s in spec["sources"]],
             extra_cflags=spec["extra_cflags"],
             build_directory=bdir, verbose=verbose, is_python_module=True)
        # keep only the shared object: objects and ninja files are build scratch
        for f in os.listdir(bdir):
            if not f.endswith(".so"):
                os.remove(os.path.join(bdir, f))
        print(f"[oracle/build_ref] built {name}")
    return True


if __name__ == "__main__":
    build(sys.argv[1:] or None, verbose=True)
