"""Recipe for `oracle/_ref/`: the UNMODIFIED reference modules of the hot path, copied verbatim from /root/reference
(this container only; the GPU box gets the already-built directory with the repo snapshot).

    python oracle/build_ref.py          # or __graft_entry__.build()

`oracle/_ref/` is git-ignored (never committed: reference sources stay out of the history) but not gpurun-ignored.
What is vendored: train.py, transformer.py, bar_distribution.py, utils.py, encoders.py, positional_encodings.py,
decoders.py and priors/{prior,utils,mlp}.py -- everything `train.train` touches that imports without gpytorch /
botorch / pyro.  The reference's `priors/__init__.py` imports all of those (priors/__init__.py:1), so the package
init is replaced by THIS repo's `oracle/ref_stub/priors_init.py`, whose `fast_gp` is the oracle's CPU restatement of
priors/fast_gp.py (dense kernel -> torch.linalg.cholesky -> matmul).  TEST / BASELINE INFRASTRUCTURE ONLY.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PFN_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")

TOP = ["train.py", "transformer.py", "bar_distribution.py", "utils.py", "encoders.py", "positional_encodings.py",
       "decoders.py"]
PRIORS = ["prior.py", "utils.py", "mlp.py"]


def build(verbose=True):
    """Returns True if oracle/_ref is (now) present."""
    if not os.path.isdir(REF):
        if verbose:
            print(f"[oracle/build_ref] {REF} not present; keeping the existing oracle/_ref ({'found' if os.path.isdir(DST) else 'absent'})")
        return os.path.isdir(DST)
    os.makedirs(os.path.join(DST, "priors"), exist_ok=True)
    for f in TOP:
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, f))
    for f in PRIORS:
        shutil.copyfile(os.path.join(REF, "priors", f), os.path.join(DST, "priors", f))
    shutil.copyfile(os.path.join(HERE, "ref_stub", "priors_init.py"), os.path.join(DST, "priors", "__init__.py"))
    shutil.copyfile(os.path.join(HERE, "ref_stub", "fast_gp_cpu.py"), os.path.join(DST, "priors", "fast_gp.py"))
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write(f"verbatim copies from {REF} made by oracle/build_ref.py: {TOP + ['priors/' + p for p in PRIORS]}\n"
                 "priors/__init__.py and priors/fast_gp.py are this repo's oracle/ref_stub files (gpytorch is not installed)\n")
    if verbose:
        print(f"[oracle/build_ref] vendored {len(TOP) + len(PRIORS)} reference modules into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
