#!/bin/bash
# A/B of the two sweep-kernel forms (separate processes: the switch is read once).
cd "$(dirname "$0")/.."
RC_SOBOL_SWEEP=park python tools/time_sobol_variants.py
python tools/time_sobol_variants.py
