#!/bin/bash
# launch table of one S frame, only the fused kernels + total
timeout 120 python tools/launch_table.py --min-ms 0 > /tmp/lt.log 2>&1; head -1 /tmp/lt.log; grep "ffn_fused\|qkv_fused" /tmp/lt.log | head -4
