"""Stage the reference's OWN hot-path source files into ``oracle/_ref/``.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference (XiaRho/CMDA) is pure Python: there is nothing to compile, and ``/root/reference`` does not
exist on the GPU box.  ``oracle/_ref/`` is git-ignored (the sources never enter this repository's history) but
NOT gpurun-ignored, so the staged files travel to the GPU box next to the built ``.so`` files, exactly like the
base contract's ``baseline/_ref`` install.  ``bench.py --impl reference`` and the ``cpu_baseline`` leg execute
the functions of these files, unmodified, through ``oracle/ref_runner.py``.

Run by ``__graft_entry__.build()`` in the build container.  The four files hold the whole path
(SURVEY.md section 8a):

    mmseg/datasets/dsec.py                     events_to_voxel_grid, events_norm, DSECDataset.get_events_vg
    mmseg/datasets/utils.py                    get_ic, get_image_change_from_pil
    create_cityscapes_image_change.py          get_image_change
    create_dsec_dataset_txt.py                 create_images_to_events_index
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("CMDA_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ("mmseg/datasets/dsec.py", "mmseg/datasets/utils.py", "create_cityscapes_image_change.py",
         "create_dsec_dataset_txt.py")


def staged() -> bool:
    return all(os.path.isfile(os.path.join(REF_DST, f)) for f in FILES)


def stage(verbose: bool = False) -> bool:
    """Copy the files when the reference tree is present; returns whether ``oracle/_ref`` is complete."""
    if not os.path.isfile(os.path.join(REF_SRC, FILES[0])):
        return staged()
    for f in FILES:
        dst = os.path.join(REF_DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, f), dst)
    with open(os.path.join(REF_DST, "README"), "w") as fh:
        fh.write("Unmodified copies of four files of XiaRho/CMDA (the reference), staged by oracle/ref_stage.py for\n"
                 "bench.py's reference arm.  Git-ignored on purpose: not part of this repository.\n")
    if verbose:
        print("staged", REF_DST)
    return staged()


if __name__ == "__main__":
    print("oracle/_ref complete:", stage(verbose=True))
