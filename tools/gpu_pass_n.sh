#!/bin/bash
timeout 300 python tools/tc_timeline.py 2>&1 | tail -14
timeout 600 python tools/debug2.py 2>&1 | tail -40
