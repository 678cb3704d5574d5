# Include from the reference's lib/CMakeLists.txt (after its list(APPEND ltetrigger_sources ...)):
#   set(LTB200_ROOT /path/to/this/repo)
#   include(${LTB200_ROOT}/gr-ltetrigger_b200/gr_oot/lib/ltetrigger_b200.cmake)
# It swaps the srsLTE-backed pss/sss implementations for the B200-backed ones; everything else
# (mib, cellstore, swig, grc, python) is built from the reference tree as before.
list(REMOVE_ITEM ltetrigger_sources pss_impl.cc sss_impl.cc)
list(APPEND ltetrigger_sources
     ${LTB200_ROOT}/gr-ltetrigger_b200/gr_oot/lib/pss_b200_impl.cc
     ${LTB200_ROOT}/gr-ltetrigger_b200/gr_oot/lib/sss_b200_impl.cc)
include_directories(${LTB200_ROOT}/include ${LTB200_ROOT}/gr-ltetrigger_b200/gr_oot/lib)
find_library(LTETRIGGER_B200_LIBRARY ltetrigger_b200 HINTS ${LTB200_ROOT}/gr-ltetrigger_b200/lib)
list(APPEND ltetrigger_libs ${LTETRIGGER_B200_LIBRARY})
