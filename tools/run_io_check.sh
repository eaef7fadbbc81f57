# full GPU suite + genome-scale files -> files time of the drop-in class with
# the staged and the plain .npy writer (outputs compared byte for byte)
timeout 170 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 150 python tools/time_class.py /tmp/h3d_class --chroms=all > gpurun_out/r02_time_class_genome.json 2> gpurun_out/r02_time_class_genome.err; echo "time_class exit $?"; tail -c 300 gpurun_out/r02_time_class_genome.err; cat gpurun_out/r02_time_class_genome.json
mv /tmp/h3d_class/out_all /tmp/h3d_class/out_staged
H3D_WRITER=plain timeout 100 python tools/time_class.py /tmp/h3d_class --chroms=all --reuse > gpurun_out/r02_time_class_genome_plain.json 2> gpurun_out/r02_time_class_genome_plain.err; echo "time_class plain exit $?"; tail -c 300 gpurun_out/r02_time_class_genome_plain.err; cat gpurun_out/r02_time_class_genome_plain.json
timeout 60 python - <<'PY'
import filecmp, os
a, b = '/tmp/h3d_class/out_staged', '/tmp/h3d_class/out_all'
fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
same = [f for f in fa if f in fb and filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False)]
print('writer check: %d files staged, %d plain, %d byte-identical; differing: %s'
      % (len(fa), len(fb), len(same), [f for f in fa if f not in same][:5]))
PY
