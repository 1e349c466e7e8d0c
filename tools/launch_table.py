#!/usr/bin/env python
"""Print one line per kernel launch from an `ncu --csv --metrics ...` log (launch list).

    python tools/launch_table.py gpurun_out/launches.csv
"""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault((r[0], r[4][:70], r[7], r[8]), {})[r[12]] = r[14]
    for (i, name, blk, grid), v in d.items():
        t = float(v.get("gpu__time_duration.sum", "0").replace(",", "")) / 1e6
        extras = " ".join(f"{k.split('.')[0].replace('sm__inst_executed_pipe_', 'pipe_').replace('smsp__', '')}={val}" for k, val in v.items()
                          if k != "gpu__time_duration.sum")
        print(f"{i:>3} {t:9.3f} ms  {name}  grid={grid} block={blk}  {extras}")


if __name__ == "__main__":
    main()
