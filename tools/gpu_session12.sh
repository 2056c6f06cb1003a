#!/bin/bash
# round 2, GPU session 12: from-proofs path, slab size x front-end placement x phase split (wall ms per call, 3 calls each)
# (record of a measurement: options of this script that measured neutral or negative -- bv_merge_rows, l2_fetch, plan_overlap,
# bv_carveout, accumulate variants 51 / 52 -- were removed from the library afterwards; DESIGN.md sections 4, 5b and 9 quote the results)
set -u
for o in "bv_prep_stream=0 bv_chunk_terms=524288 bv_phase1_rows=12" "bv_prep_stream=0 bv_chunk_terms=524288 bv_phase1_rows=15" \
         "bv_prep_stream=0 bv_chunk_terms=1048576 bv_phase1_rows=12" "bv_prep_stream=0 bv_chunk_terms=262144 bv_phase1_rows=12" \
         "bv_prep_stream=1 bv_chunk_terms=524288 bv_phase1_rows=12" "bv_prep_stream=1 bv_chunk_terms=524288 bv_phase1_rows=15" \
         "bv_prep_stream=1 bv_chunk_terms=262144 bv_phase1_rows=15" "bv_prep_stream=1 bv_chunk_terms=1048576 bv_phase1_rows=15" \
         "bv_prep_stream=1 bv_chunk_terms=524288 bv_phase1_rows=16 bv_prep_blocks=1" "bv_prep_stream=1 bv_chunk_terms=524288 bv_phase1_rows=15 bv_prep_blocks=3 bv_prep_smem_kb=40"; do
  a=""; for kv in $o; do a="$a --opt $kv"; done
  python tools/bv_timeline.py $a 2>&1 | grep "wall ms"
done
