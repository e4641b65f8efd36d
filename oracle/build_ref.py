"""Stage the UNMODIFIED reference files needed by the hot path under oracle/_ref/ — TEST / BASELINE INFRASTRUCTURE.

The reference is pure Python (no C/C++/CUDA of its own), so "building" it means making its own source files
importable where /root/reference does not exist (the GPU box): `__graft_entry__.build()` calls `stage()` in the
authoring container, which copies — byte for byte, checked by hash — `model/*.py`, `utils/*.py` and the `*.py` files
of `eval_func/` (imported at module level by `model/utils.py:14-17`) into `oracle/_ref/`.  That directory is
git-ignored (reference sources never enter the history) but travels with the gpurun snapshot like a built `.so`.

Consumers: `bench.py --impl reference` and the `cpu_baseline` leg (`kind: "reference"`), which run the reference's own
`Trainer.update_bcd` + `BCEDiceLoss` + `torch.optim.Adam` on the host through `oracle/pv_shim` (the pytorchvideo /
fvcore classes the reference imports, absent offline); `oracle/reference_loader.py` resolves the root.
Nothing under change3d_b200/ imports this.
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
SRC_DEFAULT = "/root/reference"
SUBDIRS = ("model", "utils", "eval_func")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def staged() -> bool:
    return os.path.isfile(os.path.join(REF_OUT, "model", "trainer.py"))


def stage(src: str = SRC_DEFAULT) -> bool:
    """Copy the reference's python files for the path into oracle/_ref/.  Returns False (and leaves any previously
    staged copy alone) when the reference tree is not present."""
    if not os.path.isfile(os.path.join(src, "model", "trainer.py")):
        return False
    manifest = []
    for sub in SUBDIRS:
        for root, _, files in os.walk(os.path.join(src, sub)):
            for fn in sorted(files):
                if not fn.endswith(".py"):
                    continue
                s = os.path.join(root, fn)
                d = os.path.join(REF_OUT, os.path.relpath(s, src))
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                assert _sha(s) == _sha(d)
                manifest.append(f"{_sha(d)}  {os.path.relpath(d, REF_OUT)}")
    with open(os.path.join(REF_OUT, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference tree not found; nothing staged", REF_OUT)
