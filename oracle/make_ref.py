"""Recipe for oracle/_ref/: copy the UNMODIFIED reference files this path needs out of /root/reference.

    python oracle/make_ref.py

Nothing is edited; every copy is compared byte for byte with its source.  oracle/_ref/ is listed in .gitignore (the
repo history holds no reference sources) but NOT in .gpurunignore, so the copy travels to the GPU box, where
/root/reference does not exist: `bench.py --impl reference` and the drop-in tests run the reference's own Fit /
get_all_loss_DeepF / get_Rt_loss / modelLoader from it (oracle/ref_env.py).  __graft_entry__.build() calls this
when /root/reference is mounted.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_env import REF_COPY, REF_FILES, REF_MOUNT  # noqa: E402


def make_ref(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(REF_MOUNT, "deepFEPE")):
        if verbose:
            print(f"{REF_MOUNT} is not mounted: oracle/_ref left as it is")
        return False
    n = 0
    for rel in REF_FILES:
        src, dst = os.path.join(REF_MOUNT, rel), os.path.join(REF_COPY, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            n += 1
        assert filecmp.cmp(src, dst, shallow=False), rel
    if verbose:
        print(f"oracle/_ref: {len(REF_FILES)} reference files present ({n} copied)")
    return True


if __name__ == "__main__":
    make_ref()
