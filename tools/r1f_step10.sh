#!/bin/bash
# HEAD defaults at the headline size (crash check + one bench line)
timeout 40 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline 2>&1 | tail -1 | cut -c1-220
