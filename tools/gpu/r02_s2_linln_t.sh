#!/bin/bash
PB200_LINLN_WARPS=8 timeout 120 python tools/linln_one.py 16000 512 0.0 2>&1 | tail -12
PB200_LINLN_WARPS=16 timeout 120 python tools/linln_one.py 16000 512 0.1 2>&1 | tail -12
