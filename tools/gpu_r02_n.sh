#!/bin/bash
# one B200: bench lines of the final HEAD (kernel legs only) for Catphan and the air scan
set -u
O=gpurun_out/r02n
mkdir -p $O
timeout 200 python bench.py --legs none > $O/bench_catphan_nolegs.json 2> $O/bench.err; cut -c1-250 $O/bench_catphan_nolegs.json
timeout 100 python bench.py --workload air --legs none > $O/bench_air_nolegs.json 2>> $O/bench.err; cut -c1-250 $O/bench_air_nolegs.json
