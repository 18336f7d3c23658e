#!/bin/bash
# quick GPU check: parity tests, then ms/step at 1M beads
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python scripts/stability.py 1000000 10000 ${1:-2000} fene 2>&1 | tail -4
