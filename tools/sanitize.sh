#!/bin/bash
# compute-sanitizer over the parity scenes (small: the tools slow kernels down 10-100x). Usage: tools/sanitize.sh [memcheck|racecheck|initcheck|synccheck]
# Run on a GPU box from the repo root; output in gpurun_out/sanitize_<tool>.log
tool=${1:-memcheck}
mkdir -p gpurun_out
compute-sanitizer --tool "$tool" --error-exitcode 99 --print-limit 20 \
    python -m pytest tests/test_parity_gpu.py tests/test_state_gpu.py -m gpu -x -q \
    -k "golden and (cube2k or wall2k or ragged) or cluster or one_cell or empty_and_tiny or rebinned or paired or morton or cooperative or graph_and_pdl" \
    > "gpurun_out/sanitize_$tool.log" 2>&1
rc=$?
grep -E "ERROR SUMMARY|passed|failed|Error" "gpurun_out/sanitize_$tool.log" | tail -5
exit $rc
