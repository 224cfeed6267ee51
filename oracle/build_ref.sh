#!/usr/bin/env bash
# build_ref.sh -- TEST INFRASTRUCTURE (oracle).
#
# Compiles the reference simulators FROM WHERE THEY LIE (/root/reference, read-only) into
# shared libraries under oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
# The reference's own build system is not used (it has none beyond .vscode/tasks.json).
# No reference text is written to disk: each source is streamed through `sed` (the
# mechanical edits listed below) into gcc's stdin between ref_shim_pre.h and
# ref_shim_post.h.
#
# Edits (all outside the state machine):
#   W  RandomAccessWithNOMA.c:221      nUE sweep 10000..100000 -> the single point ref_nue
#   B  RandomAccessSimulatorBeta.c:71  same sweep line
#   B  RandomAccessSimulatorBeta.c:47-57  parameter locals (B has no CLI) read ref_p_* globals
#
# Usage: oracle/build_ref.sh [reference_dir]     (default /root/reference)
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${1:-/root/reference}"
out="$here/_ref"
inc="$here/../include"
mkdir -p "$out"
CC="${CC:-gcc}"
CFLAGS="-O3 -w -fPIC -shared -std=gnu11 -I$here -I$inc"

if [ ! -f "$ref/RandomAccessWithNOMA.c" ]; then
    echo "build_ref.sh: $ref not present; keeping prebuilt oracle/_ref (if any)" >&2
    exit 0
fi

SWEEP='s/for (int n = 10000; n <= 100000; n += 10000)/for (int n = ref_nue; n <= ref_nue; n += 10000)/'

# ---- W: RandomAccessWithNOMA.c ------------------------------------------------------------
{ echo '#include "ref_shim_pre.h"'
  sed -e "$SWEEP" "$ref/RandomAccessWithNOMA.c"
  echo; echo '#define REF_VARIANT 0'; echo '#include "ref_shim_post.h"'
} | $CC $CFLAGS -x c - -o "$out/libref_w.so" -lm

# ---- B: RandomAccessSimulatorBeta.c -------------------------------------------------------
{ echo '#include "ref_shim_pre.h"'
  sed -e "$SWEEP" \
      -e 's/int nPreamble = 54; /int nPreamble = ref_p_nPreamble; /' \
      -e 's/int backoffIndicator = 20;/int backoffIndicator = ref_p_backoff;/' \
      -e 's/int nGrantUL = 54; /int nGrantUL = ref_p_nGrantUL; /' \
      -e 's/int maxRarWindow = 6;/int maxRarWindow = ref_p_maxRarWindow;/' \
      -e 's/int maxMsg2TxCount = 9;/int maxMsg2TxCount = ref_p_maxMsg2TxCount;/' \
      -e 's/int accessTime = 5; /int accessTime = ref_p_accessTime; /' \
      -e 's/int distribution = 1;/int distribution = ref_p_distribution;/' \
      "$ref/RandomAccessSimulatorBeta.c"
  echo; echo '#define REF_VARIANT 1'; echo '#include "ref_shim_post.h"'
} | $CC $CFLAGS -x c - -o "$out/libref_b.so" -lm

# ---- N: NOMA.c ---------------------------------------------------------------------------
#   N:644  seed loop 0..9 -> one seed;  N:648  nUE sweep -> ref_nue
#   N:499-546 (resourceRequestAllocation: `user` is the array) rand() -> keyed by (user+i)->idx
#   N:194-324 (preambleSectorCollisionDetection)              rand() -> base-station stream (sector s)
{ echo '#include "ref_shim_pre.h"'
  sed -e '644s/seed < 10/seed < 1/' \
      -e '648s/for (int nUE = 10000; nUE <= 100000; nUE += 10000)/for (int nUE = ref_nue; nUE <= ref_nue; nUE += 10000)/' \
      -e '499,546s/rand()/ref_tape_rand((user + i)->idx, time)/g' \
      -e '194,324s/rand()/ref_tape_rand_bs(s, time)/g' \
      "$ref/NOMA.c"
  echo; echo '#include "ref_shim_post_n.h"'
} | $CC $CFLAGS -x c - -o "$out/libref_n.so" -lm

# ---- N2: NOMA.c with its alternative, non-sector collision function switched in -----------
#   N:688-689  the commented-out call of preambleCollisionDetection (N:325-447) replaces the sector one
#   N:325-447  its rand() (N:411, `user` is the array) -> base-station stream, sector 0
{ echo '#include "ref_shim_pre.h"'
  sed -e '644s/seed < 10/seed < 1/' \
      -e '648s/for (int nUE = 10000; nUE <= 100000; nUE += 10000)/for (int nUE = ref_nue; nUE <= ref_nue; nUE += 10000)/' \
      -e '499,546s/rand()/ref_tape_rand((user + i)->idx, time)/g' \
      -e '194,324s/rand()/ref_tape_rand_bs(s, time)/g' \
      -e '325,447s/rand()/ref_tape_rand_bs(0, time)/g' \
      -e '688s#// preambleCollisionDetection#preambleCollisionDetection#' \
      -e '689s#preambleSectorCollisionDetection#// preambleSectorCollisionDetection#' \
      "$ref/NOMA.c"
  echo; echo '#include "ref_shim_post_n.h"'
} | $CC $CFLAGS -x c - -o "$out/libref_n2.so" -lm

# ---- U0: RandomAccessSimulator.c ----------------------------------------------------------
#   U0:42  nUE sweep -> ref_nue;  U0:48-49 parameter locals -> ref_p_*
#   U0:84  `i <= activeCheck` reads one element past the array once activeCheck reaches nUE -> clamp
#   U0:321 `float averageDelay = averageDelay/...` redeclares the parameter (does not compile) -> the `float` dropped
#          (plain assignment to the parameter, so that U0:322-323 report the average as the text says)
{ echo '#include "ref_shim_pre.h"'
  sed -e '42s/for(int n = 10000; n <= 100000; n+=10000)/for(int n = ref_nue; n <= ref_nue; n+=10000)/' \
      -e '48s/int nPreamble = 64;/int nPreamble = ref_p_nPreamble;/' \
      -e '49s/int backoffIndicator = 20;/int backoffIndicator = ref_p_backoff;/' \
      -e '84s/i <= activeCheck/i <= activeCheck \&\& i < nUE/' \
      -e '321s/float averageDelay = averageDelay/averageDelay = averageDelay/' \
      "$ref/RandomAccessSimulator.c"
  echo; echo '#include "ref_shim_post_u0.h"'
} | $CC $CFLAGS -x c - -o "$out/libref_u0.so" -lm

echo "built: $(ls "$out")"
