#!/usr/bin/env python
"""Summarise a parity_dump.py .npz: GPU vs oracle and oracle vs its own one-ulp-perturbed run, per exit-ray row."""
import sys
import numpy as np
d = np.load(sys.argv[1])
rf, rfo, rf1 = d['rf'], d['rf_o'], d['rf_1']
print('rays', rf.shape[1], ' steps per ray identical:', bool(np.array_equal(d['steps'], d['steps_o'])), ' mean steps', float(d['steps_o'].mean()))
scale = np.sqrt((rfo ** 2).mean(axis=1))
for name, a, b in (('GPU vs oracle', rf, rfo), ('oracle(one-ulp-moved rays) vs oracle', rf1, rfo)):
    ab = np.abs(a - b)
    print(name)
    for r, nm in enumerate(('x [m]', 'theta [rad]', 'y [m]', 'phi [rad]')):
        own = ab[r] / np.abs(b[r])
        print('  %-12s abs: max %.2e p99.9 %.2e median %.2e | rel to own value: max %.2e p99.9 %.2e median %.2e | rel to max(|v|, row rms %.2e): max %.2e'
              % (nm, ab[r].max(), np.percentile(ab[r], 99.9), np.median(ab[r]), own.max(), np.percentile(own, 99.9), np.median(own), scale[r],
                 (ab[r] / np.maximum(np.abs(b[r]), scale[r])).max()))
