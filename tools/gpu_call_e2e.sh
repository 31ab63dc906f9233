set -x
timeout -s KILL 200 python tools/pcie_probe.py 2>&1 | tail -20
