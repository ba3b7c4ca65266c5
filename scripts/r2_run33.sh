cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 2000 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tail -9
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
