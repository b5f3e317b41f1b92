tag=r02zs; out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { echo "build failed"; exit 1; }
timeout 600 python -m pytest tests -m gpu -q -k "canar or only_the_rows or scheduler or aligned or overlapped or one_prologue" > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $out/${tag}_pytest_gpu.log
