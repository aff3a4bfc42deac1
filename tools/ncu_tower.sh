set -u
O=gpurun_out/ncu_tower
mkdir -p $O
ncu --clock-control none --set full -k regex:conv_tower -s 4 -c 1 -f -o /tmp/tower python tools/probe_net.py 5 128 4096 predict > $O/tower.log 2>&1
ncu -i /tmp/tower.ncu-rep --page raw --csv > $O/tower.raw.csv 2>/dev/null
ncu -i /tmp/tower.ncu-rep --page details > $O/tower.details.txt 2>/dev/null
