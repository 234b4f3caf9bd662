#!/bin/bash
# PCIe floor vs the e2e leg (see tools/pcie_probe.py)
mkdir -p gpurun_out
timeout 600 python tools/pcie_probe.py 400000 > gpurun_out/pcie_probe.log 2>&1
tail -8 gpurun_out/pcie_probe.log
