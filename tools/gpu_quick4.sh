#!/bin/bash
for d in 0 1 2 4 7; do echo "HIG_APPLY_DBG=$d"; HIG_APPLY_DBG=$d python tools/step_breakdown.py 100 2>&1 | grep "attn apply + stylize"; done
