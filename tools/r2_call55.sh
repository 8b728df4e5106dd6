#!/usr/bin/env bash
mkdir -p gpurun_out
{
timeout 600 python tools/profile_levels.py 8:265 8:18 18:33 33:48 48:63 63:86 86:101 101:143 143:163 163:186 186:205 205:225 225:245 245:265 48:205
} > gpurun_out/call55.log 2>&1
