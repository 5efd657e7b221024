#!/bin/bash
for d in 0 1 2 3; do echo "== HIG_AP_DBG=$d"; HIG_AP_DBG=$d timeout 200 python tools/step_breakdown.py 100 2>&1 | grep "attn apply\|full denoiser"; done
