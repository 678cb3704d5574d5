#!/bin/sh
# The four command lines of the reference's examples/test.sh on the bundled test frames (needs a
# B200; tests/test_gpu_blocks.py::test_cell_search_file_cli asserts the same runs), then the batched
# forms this repository adds.
set -o verbose
cd "$(dirname "$0")"
F=../tests/golden/test_frames

./cell_search_file.py --sample-rate 1.92M $F/lte_frame_6prb_cellid_123 --repeat --time-out 1
./cell_search_file.py --sample-rate 7.68M $F/lte_frame_25prb_cellid_124 --repeat --time-out 1
./cell_search_file.py --sample-rate 15.36M $F/lte_frame_50prb_cellid_125 --repeat --time-out 1
./cell_search_file.py --sample-rate 30.72M $F/lte_frame_100prb_cellid_369 --repeat --time-out 1

./cell_search_batch.py --sample-rate 7.68M $F/lte_frame_25prb_cellid_124 $F/lte_frame_25prb_cellid_124 --repeat --cut-off 7.68M
./snr_sweep.py --streams 64 --seconds 0.5 --snr-min -6 --snr-max 6 --snr-step 3
