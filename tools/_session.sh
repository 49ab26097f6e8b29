#!/bin/bash
set -u
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
V="python tools/kernel_variants.py"
( $V 10000000 10000 20000; $V 150000000 10000 20000; K=31 $V 400000000 10000 20000; $V 50000000 2000000 150 fixed ) 2>&1 | grep -v "^Found\|^Done\|^Reading" | sed -e 's/ lf=0.5 buckets=[0-9]*//' -e 's/UNROLL=[0-9]* //' -e 's/ L2G=32: / /' -e 's/table_MB=[0-9x]* //' -e 's/(build.*//' -e 's/filter=\(2^[0-9]*\) pair=1 mini=1 bits=2 //'
