# round-2 tuning sweep (profiling experiments, not a bench), 1 GB k=21: leaf kernel variants
for v in "" _B _C; do
echo "== lib$v"; KMER_CUDA_LIB=kmer-extension_b200/libkmer_cuda$v.so python tools/part_experiment.py 1000000 | tail -1
done
for v in "" _B; do
echo "== k=31 lib$v"; KMER_K=31 KMER_CUDA_LIB=kmer-extension_b200/libkmer_cuda$v.so python tools/part_experiment.py 1000000 | tail -1
done
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
