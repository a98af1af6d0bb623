#!/bin/bash
# quick GPU sanity: unit tests, then (optionally) a short bench
set -x
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -40
