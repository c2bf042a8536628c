cd /root/repo; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -40
