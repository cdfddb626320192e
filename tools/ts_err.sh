for env in "A=1" "TOME_ATTN_SHORT=0" "TOME_FUSED_BLOCKS=0" "TOME_ATTN_SHORT=0 TOME_FUSED_BLOCKS=0"; do
  echo "== $env"; env $env python -m pytest tests/test_fullsize_parity.py -q -m gpu -k "bf16_teacher and timesformer" -s 2>&1 | grep -o "'err': [0-9.e-]*, 'unpatched_err': [0-9.e-]*" 
done
