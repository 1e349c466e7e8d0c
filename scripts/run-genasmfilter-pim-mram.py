#!/usr/bin/env python
"""Drop-in for the reference's aim-genasm/GenASM/DPU-MRAM-filter/run-genasmfilter-pim-mram.py (same options), backed by the B200 host."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from aim_b200.run_pim import main_genasm  # noqa: E402

if __name__ == "__main__":
    sys.exit(main_genasm("filter", "mram"))
