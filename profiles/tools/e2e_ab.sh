#!/bin/bash
# same box: default bench, bench without the extras, default bench again -- does anything the extras do change the host-buffer numbers?
for tag in full1 noextras full2; do
  if [ $tag = noextras ]; then fl="--no-extras --ingest-reads 0 --no-parity-check"; else fl=""; fi
  python bench.py $fl > gpurun_out/ab_$tag.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/ab_$tag.json'));e=d['e2e'];print('$tag', 'one call %.3f  two seams %.3f  plain %.3f ms' % (e['ms_per_step'], e['two_seam_calls']['ms_per_step'], e['uncompressed']['ms_per_step']))"
done
