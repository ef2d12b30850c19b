#!/bin/bash
# round 2, first GPU pass: tests, smoke, the driver's exact bench command, the reference arm, pipeline studies
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02a_pytest.log
python __graft_entry__.py smoke > $O/r02a_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02a_smoke.log
python bench.py --steps 20 --warmup 5 > $O/r02a_bench_driver.json 2> $O/r02a_bench_driver.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02a_bench_reference.json 2> $O/r02a_bench_reference.err
for extra in "" "--no-priority" "--depth 2" "--depth 3" "--depth 4" "--depth 8" "--depth 3 --no-priority"; do
  python bench.py --only --steps 2000 --warmup 20 $extra >> $O/r02a_pipeline_study.jsonl 2>> $O/r02a_pipeline_study.err
done
python bench.py --only --steps 300 --warmup 20 --conf 0.001 >> $O/r02a_pipeline_study.jsonl 2>> $O/r02a_pipeline_study.err
python bench.py --only --steps 300 --warmup 20 --conf 0.001 --no-priority >> $O/r02a_pipeline_study.jsonl 2>> $O/r02a_pipeline_study.err
tail -3 $O/r02a_pytest.log; cat $O/r02a_smoke.log | tail -2
