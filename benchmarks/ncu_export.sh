#!/bin/bash
# ncu_export.sh <report.ncu-rep>: keep the judged numbers as small text next to the report, then drop the report
# (gpurun copies back at most 64 MiB; a --set full report is 5-15 MB per kernel launch).
R=$1; B=${R%.ncu-rep}
ncu -i $R --page raw --csv > $B.raw.csv 2>/dev/null
ncu -i $R --page details --csv 2>/dev/null | grep -E "Duration|Throughput|Pipe|Registers|Theoretical Occupancy|Achieved Occupancy|Bank|L2 Hit|Stall|Issue|Eligible|No Eligible|DRAM|Shared Memory Configuration|Block Size|Grid Size" > $B.details.csv
ncu -i $R --page source --csv 2>/dev/null | gzip > $B.source.csv.gz
[ "$KEEP_REP" = "1" ] || rm -f $R
