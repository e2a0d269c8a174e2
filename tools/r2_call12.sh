#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python tools/margins.py 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py > /tmp/b.json 2>/tmp/b.err ) 2>&1 | grep real; cut -c1-200 /tmp/b.json
