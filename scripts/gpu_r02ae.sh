#!/bin/bash
# round 2, call AE: whole -m gpu suite + smoke() on the default build
O=gpurun_out/r02ae; mkdir -p $O
# (compute-sanitizer is closed on this pool: the memcheck / racecheck lines that stood here answered rc 86 without running)
timeout 700 python -m pytest tests -m gpu -x -q --timeout 120 > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/pytest.log
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -3
