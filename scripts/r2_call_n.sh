#!/bin/bash
# round-2 GPU call N (1 GPU): what a slab with neighbours costs on ONE GPU (kernel mode, edge runs)
cd "$(dirname "$0")/.."
for ny in 1024 4096; do
echo "== Ny=$ny"
BENCH_EXTRA="--ny $ny" scripts/bench_variants.sh main 2>/dev/null | sed 's/^/plain          /'
FV2D_FORCE_MODE=1 BENCH_EXTRA="--ny $ny" scripts/bench_variants.sh main 2>/dev/null | sed 's/^/multi          /'
FV2D_FORCE_MODE=2 BENCH_EXTRA="--ny $ny" scripts/bench_variants.sh main 2>/dev/null | sed 's/^/general        /'
FV2D_FORCE_EDGE_RUNS=1 BENCH_EXTRA="--ny $ny" scripts/bench_variants.sh main 2>/dev/null | sed 's/^/plain+edgeruns /'
FV2D_FORCE_MODE=1 FV2D_FORCE_EDGE_RUNS=1 BENCH_EXTRA="--ny $ny" scripts/bench_variants.sh main 2>/dev/null | sed 's/^/multi+edgeruns /'
done
