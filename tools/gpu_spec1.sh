#!/bin/bash
# one-GPU timing of the global-spectrum local count (one large group), plus 2 chunked rounds
python tools/spectrum_ngpu.py 2000 0 1 2>&1 | tail -1
python tools/spectrum_ngpu.py 20000 0 0 2>&1 | tail -1
python tools/spectrum_ngpu.py 30000 0 0 2>&1 | tail -1
python tools/spectrum_ngpu.py 40000 1000000 0 2>&1 | tail -1
