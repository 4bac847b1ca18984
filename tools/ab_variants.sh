#!/bin/bash
# A/B timing of engine builds: tools/ab_variants.sh base variants/lib_A.so ...   (RGPOT_B200_ENGINE selects the library)
for lib in "$@"; do
  if [ "$lib" = "base" ]; then unset RGPOT_B200_ENGINE; else export RGPOT_B200_ENGINE="$PWD/$lib"; fi
  echo "== $lib"
  python tools/bench_tile.py --variants "tile=0" --steps 8 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('  %-5s %.4f ms  %.4e atom-evals/s  E=%.12f' % (d['what'], d['ms'], d['atom_evals_per_s'], d['energy']))"
  python tools/bench_batch_dev.py --configs 32768 --steps 4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('  batch %.4f ms  %.4e atom-evals/s  Esum=%.9f' % (d['ms'], d['atom_evals_per_s'], d['e_sum']))"
done
