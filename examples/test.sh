#!/bin/sh
# What the reference's examples/test.sh exercises (its CLI on each bundled test frame at that frame's
# sample rate, looping the file for up to a second), then the batched tools this repository adds.
# Needs a B200; tests/test_gpu_blocks.py::test_cell_search_file_cli asserts the same runs.
cd "$(dirname "$0")" || exit 1
frames=../tests/golden/test_frames
set -x
for pair in 1.92M:6prb_cellid_123 7.68M:25prb_cellid_124 15.36M:50prb_cellid_125 30.72M:100prb_cellid_369; do
  python cell_search_file.py -s "${pair%%:*}" "$frames/lte_frame_${pair#*:}" --repeat --time-out 1
done
python cell_search_batch.py -s 7.68M "$frames/lte_frame_25prb_cellid_124" "$frames/lte_frame_25prb_cellid_124" --repeat --cut-off 7.68M
python snr_sweep.py --streams 64 --seconds 0.5 --snr-min -6 --snr-max 6 --snr-step 3
