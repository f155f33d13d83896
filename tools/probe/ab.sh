python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== atlas"; python tools/variant_bench.py libpmvs_b200.so 2>&1 | tail -1
echo "== no atlas"; PMVSB_NO_ATLAS=1 python tools/variant_bench.py libpmvs_b200.so 2>&1 | tail -1
