python -m pytest tests -m gpu -q -x 2>&1 | tail -5
bash profiles/tools/n2probe.sh 2>&1 | grep -v DEVICE
python profiles/tools/phase_timeline.py per256 65536 2>&1 | grep -v DEVICE | head -14
