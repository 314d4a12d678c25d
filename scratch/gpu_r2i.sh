#!/bin/bash
python -m pytest tests -m gpu -q -k "config3 or config4 or streaming or wavefront" 2>&1 | tail -4
python scratch/ab_quant.py
