#!/bin/bash
# A/B of the sweep-kernel forms and of the column tiles per CTA (separate processes: the switches are read once).
cd "$(dirname "$0")/.."
RC_SOBOL_SWEEP=park python tools/time_sobol_variants.py | cut -c1-60
for c in 1 2 4 8 16 64; do echo chunk $c; RC_SOBOL_CHUNK=$c python tools/time_sobol_variants.py | cut -c1-60; done
