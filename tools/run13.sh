#!/bin/bash
# GPU box: full validation of the final round-2 code (tests, smoke, every bench line)
cd "$(dirname "$0")/.."
bash tools/final_check.sh r02h
