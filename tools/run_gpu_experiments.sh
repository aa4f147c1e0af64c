for w in c3 c4 c2; do timeout 900 python tools/fullsize_parity.py $w 384 2>&1 | grep -v Warning | tee -a gpurun_out/fullsize_parity.log; done
