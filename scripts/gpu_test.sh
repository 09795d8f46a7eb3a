#!/bin/bash
# parity tests only. Usage: gpu_test.sh [pytest args]
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 "$@" 2>&1 | tail -15
