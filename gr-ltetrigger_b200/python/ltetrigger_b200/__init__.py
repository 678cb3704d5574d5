"""ltetrigger_b200 -- host-side mirror of the reference's Python surface for the PSS+SSS
search path, on top of the sm_100a C-ABI library.

Reference surface mirrored here (python/__init__.py:28-31 of the reference exports
`pss`, `sss`, `mib`, `cellstore`, `downlink_trigger_c`; the first two and the hier block
are this path):
    ltetrigger.pss(N_id_2, psr_threshold, track_after=16, track_every=8)
    ltetrigger.sss(N_id_2)
    ltetrigger.downlink_trigger_c(psr_threshold, exit_on_success=False)
plus the batched `Trigger` engine the blocks are built on.
"""
from ._abi import (LIB_PATH, SUCCESS, ERROR, ERROR_INVALID_INPUTS, SLOT_LEN, HALF_FRAME, SYMBOL_SZ, CONV_LEN,
                   LOOKAHEAD, FMT_FC32, FMT_SC16, FMT_SC8, MAX_DECIM, CORR_DIRECT, CORR_FFT, OS_STEP, FRAME_FDD, FRAME_TDD, FRONTEND_FP32, FRONTEND_TC_INT, PIPE_OVERLAP, PIPE_SERIAL, MIN_PSR_THRESHOLD, F_SEARCHED, F_OVER, F_EMIT, F_TRACKING,
                   F_TAG_LOST, F_SSS, F_CELL, F_CP_NORM, WINDOW_REC, LtbError, lib)
from .engine import Trigger, device_count, kernel_pss_corr, kernel_pss_corr_fft, kernel_decimate, kernel_decimate_tc, tables
from .blocks import pss, sss, mib, cellstore, downlink_trigger_c, tag_t
