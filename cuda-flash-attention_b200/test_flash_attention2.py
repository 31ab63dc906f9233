#!/usr/bin/env python
"""Entry point with the reference harness's name and flags: python test_flash_attention2.py --mode both ...
(implementation: fa2_b200/harness.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fa2_b200.harness import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main())
