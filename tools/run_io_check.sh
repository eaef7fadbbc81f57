# full GPU suite + genome-scale files -> files time of the drop-in class
timeout 170 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 150 python tools/time_class.py /tmp/h3d_class --chroms=all > gpurun_out/r02_time_class_genome.json 2> gpurun_out/r02_time_class_genome.err; echo "time_class exit $?"; tail -c 400 gpurun_out/r02_time_class_genome.err; cat gpurun_out/r02_time_class_genome.json
