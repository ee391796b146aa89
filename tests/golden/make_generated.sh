#!/bin/bash
# Regenerates tests/golden/generated.9.6 with the reference's own writer (tools/generateMatrix.cpp, built by
# `make -C oracle ref`).  The tool seeds rand() with time(NULL), so every run gives different values; the committed
# file is one such run and the tests only rely on the format and on round-tripping its values.
set -e
make -s -C "$(dirname "$0")/../../oracle" ref
"$(dirname "$0")/../../oracle/_ref/generateMatrix" 9 6 > "$(dirname "$0")/generated.9.6"
