"""Stage the UNMODIFIED reference package next to the oracle so that it travels to the GPU box (TEST INFRASTRUCTURE).

    python -m oracle.stage_ref          # build container only: needs /root/reference

Copies `/root/reference/coperception/coperception/**/*.py` (57 files, the pure-Python package; the cp37 `mapping*.so` is
skipped) to `oracle/_ref/coperception/` and `tools/det/{train,test}_codet.py` to `oracle/_ref/tools/det/`.  `oracle/_ref/` is git-ignored -- reference sources never enter the history --
but NOT gpurun-ignored, so `pytest -m gpu` on the B200 box can import the reference's own `FaFModule.step` /
`predict_all` / `DiscoNet` (tests/test_callers_gpu.py) and run them against the drop-in classes.
`__graft_entry__.build()` calls `stage()` whenever /root/reference is present.
"""
import os
import shutil

SRC = "/root/reference/coperception/coperception"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "coperception")


def stage(verbose: bool = False) -> bool:
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    n = 0
    for root, _, files in os.walk(SRC):
        rel = os.path.relpath(root, SRC)
        for f in files:
            if not f.endswith(".py"):
                continue
            out_dir = os.path.join(DST, rel) if rel != "." else DST
            os.makedirs(out_dir, exist_ok=True)
            src, dst = os.path.join(root, f), os.path.join(out_dir, f)
            if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src) or os.path.getsize(dst) != os.path.getsize(src):
                shutil.copyfile(src, dst)
            n += 1
    # the two unmodified tool scripts the drop-in has to serve (north_star): tools/det/train_codet.py, test_codet.py
    tsrc, tdst = os.path.join(os.path.dirname(SRC), "tools", "det"), os.path.join(os.path.dirname(DST), "tools", "det")
    os.makedirs(tdst, exist_ok=True)
    for f in ("train_codet.py", "test_codet.py"):
        if os.path.exists(os.path.join(tsrc, f)):
            shutil.copyfile(os.path.join(tsrc, f), os.path.join(tdst, f))
            n += 1
    if verbose:
        print(f"staged {n} reference files -> {DST}")
    return True


def tool_path(name: str) -> str:
    """Path of an unmodified reference tool script (live tree when present, else the staged copy)."""
    live = os.path.join(os.path.dirname(SRC), "tools", "det", name)
    return live if os.path.exists(live) else os.path.join(os.path.dirname(DST), "tools", "det", name)


if __name__ == "__main__":
    stage(verbose=True)
