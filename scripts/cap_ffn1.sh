O=gpurun_out; mkdir -p $O
export_rep() { ncu -i $1.ncu-rep --page details > $1_ncu_details.txt 2>&1; ncu -i $1.ncu-rep --page raw --csv > $1_ncu_raw.csv 2>&1;
               ncu -i $1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $1_ncu_source.csv.gz; rm -f $1.ncu-rep; }
python scripts/one_gemm.py 41472 2048 512 256 129 > $O/plain_g9.log 2>&1; cat $O/plain_g9.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persist -s 3 -c 1 -o $O/r02_ffn1_gelugrad_pair_v8 -f python scripts/one_gemm.py 41472 2048 512 256 129 > $O/ncu_g9.log 2>&1
echo "exit $?"; export_rep $O/r02_ffn1_gelugrad_pair_v8
python scripts/one_gemm.py 30000 2048 512 256 1 > $O/plain_g3.log 2>&1; cat $O/plain_g3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persist -s 3 -c 1 -o $O/r02_ffn1_gelu_decode_v8 -f python scripts/one_gemm.py 30000 2048 512 256 1 > $O/ncu_g3.log 2>&1
echo "exit $?"; export_rep $O/r02_ffn1_gelu_decode_v8
python scripts/one_gemm.py 41472 2048 512 256 0 > $O/plain_g0.log 2>&1; cat $O/plain_g0.log
python scripts/one_attn.py > $O/plain_attn.log 2>&1; tail -5 $O/plain_attn.log
