/* The C ABI from plain C (C99, no C++): the header compiles, the library links, the host-only
 * entry points work, and on a machine with a B200 one fixture goes through the engine.
 *   usage: test_abi [fc32 file at 1.92 Msps]                                              */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ltetrigger_b200.h"

int main(int argc, char **argv)
{
  float hr[128], hi[128], taps[4096];
  ltb_trigger_config cfg;
  ltb_trigger *t = NULL;
  int rc, ntaps;

  printf("%s\n", ltb_version());
  if (ltb_table_pss_taps(1, hr, hi) != LTB_SUCCESS) return 1;
  ntaps = ltb_table_decim_taps(16, taps, 4096);
  printf("taps %d h1[0] %.9g\n", ntaps, hr[0]);
  if (ntaps != 525) return 1;

  memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = sizeof cfg;
  cfg.n_streams = 1;
  cfg.decim = 1;
  cfg.max_chunk = 96000;
  cfg.psr_threshold = 4.0f;
  cfg.record_all = 1;
  cfg.corr_mode = LTB_CORR_FFT;
  rc = ltb_trigger_create(&cfg, &t);
  if (ltb_device_count() == 0) {
    printf("no device: create -> %d (%s)\n", rc, ltb_last_error());
    return rc == LTB_ERROR ? 0 : 1;          /* fails loudly, no CPU path */
  }
  if (rc != LTB_SUCCESS) { printf("create failed: %s\n", ltb_last_error()); return 1; }
  if (argc > 1) {
    FILE *f = fopen(argv[1], "rb");
    static ltb_cf frame[19200], chunk[96000];
    static ltb_window_rec recs[64];
    int i, pass, n_recs = 0, cells = 0, cell_id = -1;
    if (!f || fread(frame, sizeof(ltb_cf), 19200, f) != 19200) { printf("cannot read %s\n", argv[1]); return 1; }
    fclose(f);
    for (i = 0; i < 96000; i++) chunk[i] = frame[i % 19200];
    for (pass = 0; pass < 6; pass++) {
      if (ltb_trigger_process_host(t, chunk, 0, 96000, recs, 64, &n_recs) != LTB_SUCCESS) { printf("process: %s\n", ltb_last_error()); return 1; }
      for (i = 0; i < n_recs; i++)
        if (recs[i].flags & LTB_F_CELL) { cells++; cell_id = recs[i].cell_id; }
    }
    printf("cells %d cell_id %d\n", cells, cell_id);
    if (cells == 0) return 1;
  }
  ltb_trigger_destroy(t);
  return 0;
}
