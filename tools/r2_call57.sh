#!/usr/bin/env bash
mkdir -p gpurun_out
{
timeout 600 python tools/profile_levels.py 87:88 88:89 87:89 102:103 103:104 129:130 64:65 65:66 64:66 76:77 49:50 53:54 54:55 52:53 100:101 121:122 142:143 8:9 23:24 24:25 259:260 264:265
} > gpurun_out/call57.log 2>&1
