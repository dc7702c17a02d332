#!/usr/bin/env python
"""Print the metric table of `ncu --page details --csv` output(s), one line per metric (rules dropped)."""
import csv
import sys

for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    print(path, "|", rows[1][ix["Kernel Name"]], rows[1][ix["Grid Size"]], rows[1][ix["Block Size"]])
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"] or not r[ix["Metric Name"]]:
            continue
        print(f"  {r[ix['Section Name']][:34]:34s} | {r[ix['Metric Name']]:48s} | {r[ix['Metric Unit']]:12s} | {r[ix['Metric Value']]}")
