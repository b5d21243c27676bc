#!/bin/bash
# development aid: the contract number for 1..4 parallel branches, then the default line with all extras
for s in 1 2 3 4; do python bench.py --streams $s --no-extras --steps 2000 --warmup 20; done
python bench.py --phases
