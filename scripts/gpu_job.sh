set -x
python -m pytest tests -m gpu -q -k "not tensor_core" 2>&1 | tail -5
echo "=== TC cta_group 1"
CFN_TC_CTA_GROUP=1 timeout -s KILL 300 python -m pytest tests -m gpu -q -k "tensor_core" 2>&1 | tail -40
echo "=== TC cta_group 2"
CFN_TC_CTA_GROUP=2 timeout -s KILL 300 python -m pytest tests -m gpu -q -k "tensor_core" 2>&1 | tail -40
