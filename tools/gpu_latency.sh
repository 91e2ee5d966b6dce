#!/bin/bash
# single-frame latency (720p, cuAprilTagsDetect) and the launch list of its last frame.  $1 = tag
TAG=${1:-lat}
mkdir -p gpurun_out
python tools/latency_probe.py 200 | tee gpurun_out/${TAG}_latency.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_lat_launches_raw.csv python tools/latency_probe.py 6 > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${TAG}_lat_launches_raw.csv > gpurun_out/${TAG}_lat_launches.csv; cat gpurun_out/${TAG}_lat_launches.csv
