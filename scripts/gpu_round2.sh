#!/bin/bash
# Round-2 GPU pass: tests, smoke, bench, ncu launch lists (profiler windows) and --set full captures. Logs to gpurun_out/.
# usage: scripts/gpu_round2.sh [tag] [steps...]   steps: tests smoke bench lists full   (default: all)
TAG=${1:-a}; shift
STEPS=${*:-tests smoke bench lists full}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu_$TAG.txt 2>&1
has() { [[ " $STEPS " == *" $1 "* ]]; }
NCU_LIST="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
NCU_FULL="ncu --set full --clock-control none --import-source on --profile-from-start off"
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider --timeout 600 -s > $O/r02_pytest_gpu_$TAG.log 2>&1
  echo "pytest exit $?"; grep -E "passed|failed|worst|rel err" $O/r02_pytest_gpu_$TAG.log | tail -15
fi
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -3 $O/r02_smoke_$TAG.log
fi
if has bench; then
  timeout 900 python bench.py > $O/r02_bench_$TAG.json 2> $O/r02_bench_$TAG.err; echo "bench exit $?"; tail -c 600 $O/r02_bench_$TAG.err; head -c 1500 $O/r02_bench_$TAG.json
fi
if has lists; then
  GCT_PROFILE_B=30000 timeout 600 python scripts/profile_decode.py > $O/plain_decode.log 2>&1 && \
  GCT_PROFILE_B=30000 timeout 900 $NCU_LIST --log-file $O/r02_decode_b30000_steps48-51_launches_$TAG.csv python scripts/profile_decode.py > $O/ncu_decode.log 2>&1
  echo "decode list exit $?"
  timeout 600 python scripts/profile_train.py > $O/plain_train.log 2>&1 && \
  timeout 900 $NCU_LIST --log-file $O/r02_train_step_cfg3_launches_$TAG.csv python scripts/profile_train.py > $O/ncu_train.log 2>&1
  echo "train list exit $?"
  python scripts/summarize_launches.py $O/r02_decode_b30000_steps48-51_launches_$TAG.csv > $O/r02_decode_b30000_summary_$TAG.txt 2>&1
  python scripts/summarize_launches.py $O/r02_train_step_cfg3_launches_$TAG.csv > $O/r02_train_summary_$TAG.txt 2>&1
  head -30 $O/r02_decode_b30000_summary_$TAG.txt; head -40 $O/r02_train_summary_$TAG.txt
fi
if has full; then
  # the roofline kernel at the benched batch (traffic figure of bench.py's roofline line) and the latent-space cross-attention.
  # The .ncu-rep files (tens of MB with the imported source) stay on the box: only their text pages travel back.
  export_rep() { ncu -i $1.ncu-rep --page details > $1_ncu_details.txt 2>&1; ncu -i $1.ncu-rep --page raw --csv > $1_ncu_raw.csv 2>&1;
                 ncu -i $1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $1_ncu_source.csv.gz; rm -f $1.ncu-rep; }
  timeout 300 python scripts/one_decode_attn.py 30000 49 > $O/plain_da.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_attn -s 3 -c 1 -o $O/r02_decode_attn_b30000_t49_$TAG -f \
      python scripts/one_decode_attn.py 30000 49 > $O/ncu_da.log 2>&1
  echo "decode_attn full exit $?"; export_rep $O/r02_decode_attn_b30000_t49_$TAG
  GCT_PROFILE_B=30000 timeout 900 $NCU_FULL -k regex:decode_zattn_kernel -c 1 -o $O/r02_decode_zattn_b30000_$TAG -f python scripts/profile_decode.py > $O/ncu_za.log 2>&1
  echo "zattn full exit $?"; export_rep $O/r02_decode_zattn_b30000_$TAG
  if has heavy; then
  timeout 900 $NCU_FULL -k regex:norm_bwd_kernel -s 4 -c 1 -o $O/r02_norm_bwd_$TAG -f python scripts/profile_train.py > $O/ncu_nb.log 2>&1
  echo "norm_bwd full exit $?"; export_rep $O/r02_norm_bwd_$TAG
  timeout 900 $NCU_FULL -k regex:attn_bwd_tc_kernel -s 4 -c 1 -o $O/r02_attn_bwd_$TAG -f python scripts/profile_train.py > $O/ncu_ab.log 2>&1
  echo "attn_bwd full exit $?"; export_rep $O/r02_attn_bwd_$TAG
  fi
  grep -E "dram__bytes_(read|write).sum|gpu__time_duration.sum" $O/*_ncu_raw.csv | head -5
fi
ls -la $O | tail -30
