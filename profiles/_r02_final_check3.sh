# round 2, last check of the committed build: the whole gpu tier and smoke
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02f_tests_final.log 2>&1
grep -E "passed|failed" gpurun_out/r02f_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
