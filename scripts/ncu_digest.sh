#!/bin/bash
# ncu_digest.sh REPORT.ncu-rep OUT_PREFIX : text digests of a full ncu capture (raw-page summary + per-basic-block source
# page), written next to the report so that only text has to travel back from the GPU box.
set -e
rep=$1; out=$2
ncu -i "$rep" --page raw --csv > "$out.raw.csv" 2>/dev/null
python scripts/ncu_summary.py "$out.raw.csv" > "$out.summary.txt"
python scripts/ncu_blocks.py "$rep" > "$out.blocks.txt" 2>/dev/null || true
