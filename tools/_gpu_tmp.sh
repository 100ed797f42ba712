cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "quads x|passed|failed|error|Error|peer-memory|tests/" | tail -15
