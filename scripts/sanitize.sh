#!/usr/bin/env bash
# compute-sanitizer over the CUDA path: memcheck, racecheck and synccheck on the golden-vector test, the LDE and
# MMCS tests, a full recursion-layer proof, the aggregation-tree test and the table-fill tests. Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh gpurun_out/r2_sanitize'
# Writes <prefix>_<tool>.log per tool and <prefix>_summary.txt (one line per tool: pytest verdict + sanitizer error count).
# The sanitizer slows kernels 10-100x, so the selection keeps to the small / medium cases (KoalaBear parameter set only).
set -u
prefix=${1:-gpurun_out/sanitize}
mkdir -p "$(dirname "$prefix")"
sel='tests/test_golden.py::test_cuda_path_reproduces_golden_vectors tests/test_gpu_parity.py::test_coset_lde tests/test_gpu_parity.py::test_mmcs_commit_mixed_heights tests/test_gpu_parity.py::test_recursion_layer_tables_bit_identical tests/test_gpu_parity.py::test_gpu_alu_table_fill_matches_reference_builder tests/test_gpu_parity.py::test_gpu_poseidon2_table_fill_matches_reference_builder tests/test_gpu_parity.py::test_work_queue_row_hashing_matches_one_cta_per_rows tests/test_gpu_tree.py::test_write_rows_matches_a_fresh_upload'
: > "${prefix}_summary.txt"
for tool in memcheck racecheck synccheck; do
    log="${prefix}_${tool}.log"
    timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 86 \
        python -m pytest $sel -m gpu -x -q -k "koala or not baby" > "$log" 2>&1
    rc=$?
    verdict=$(grep -E "[0-9]+ (passed|failed)" "$log" | tail -1)
    errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    echo "$tool: rc=$rc | pytest: ${verdict:-none} | ${errs:-no summary line}" | tee -a "${prefix}_summary.txt"
done
