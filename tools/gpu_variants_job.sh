#!/bin/bash
# One gpurun call: kernel time of every library variant under tools/variants/ (tools/variants.sh) on
# configs[1] and on the maize-scale genome, then a short parity subset with the default library.
#   gpurun --timeout 900 -- 'bash tools/gpu_variants_job.sh'
mkdir -p gpurun_out
{
for w in arabidopsis; do
  for v in tools/variants/*.so; do CROPSR_B200_LIB=$v timeout 300 python tools/variant_scan.py $w 30; done
done
for v in ${BIG_VARIANTS:-tools/variants/*.so}; do CROPSR_B200_LIB=$v timeout 400 python tools/variant_scan.py maize 8; done
} 2>&1 | grep -v "^$" | tee gpurun_out/variants.txt
CROPSR_B200_LIB=${PYTEST_LIB:-} timeout 600 python -m pytest tests -x -q -m gpu -k "${PYTEST_K:-checked_build or whole_candidate_table_of_configs1 or random_fastas or tile_boundaries or extras_and_annotation}" 2>&1 | tail -5 | tee gpurun_out/variants_pytest.txt
