#!/usr/bin/env python
"""`python run.py {check,train,validate,predict} ...` — same sub-commands and flags as the reference's run.py, wired
to the B200 hot path (see prostate-cancer-multimodal-segmentation_b200/cli.py)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
if __name__ == "__main__":
    importlib.import_module("prostate-cancer-multimodal-segmentation_b200.cli").main()
