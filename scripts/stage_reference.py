#!/usr/bin/env python
"""Stage the reference's main.py under baseline/_ref/wis_reference/ (git-ignored, travels to the GPU box with gpurun) so that
tests/test_dropin_reference.py can run the reference's own do_whisper / do_translate against the real engine on the GPU.
Nothing from the reference enters the repository history: baseline/_ref/ is in .gitignore."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/main.py"
DST = os.path.join(ROOT, "baseline", "_ref", "wis_reference")

if not os.path.isfile(SRC):
    sys.exit("no /root/reference on this box")
os.makedirs(DST, exist_ok=True)
shutil.copyfile(SRC, os.path.join(DST, "main.py"))
print("staged", os.path.join(DST, "main.py"))
