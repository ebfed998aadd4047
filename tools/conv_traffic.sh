# usage: bash tools/conv_traffic.sh <tag>   DRAM traffic of every conv_tc launch of one 8-receipt pass (ncu, two metrics);
# tools/conv_traffic_json.py turns the csv into profiles/<tag>_conv_traffic.json, which bench.py reports per launch
TAG=${1:-r01f}
python tools/prof_kernels.py mem > gpurun_out/traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc --csv --log-file gpurun_out/conv_traffic_${TAG}.csv python tools/prof_kernels.py mem > gpurun_out/traffic_ncu.log 2>&1
wc -l gpurun_out/conv_traffic_${TAG}.csv
