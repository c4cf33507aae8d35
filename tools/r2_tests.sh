cd /root/repo
python -m pytest ${TESTS:-tests} -m gpu -x -q 2>&1 | tail -15
