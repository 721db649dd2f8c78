#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 200 python tools/graph_probe.py 2>&1 | tail -11
