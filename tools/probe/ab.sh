# A/B of the hot kernel on one box: previous build, current build, current build without the locality order / without the atlas
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/variant_bench.py prev.so libpmvs_b200.so 2>&1 | tail -2
echo "== PMVSB_NO_ORDER=1"; PMVSB_NO_ORDER=1 python tools/variant_bench.py libpmvs_b200.so 2>&1 | tail -1
echo "== 1M patches"; python tools/variant_bench.py --patches 1048576 --reps 2 prev.so libpmvs_b200.so 2>&1 | tail -2
