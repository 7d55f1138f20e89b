"""Stage 2 -- combine the per-file count tables and normalise.  Mirror of expectedCombination.py.

    main(outputDirectory, storedExpInput, fileTag, verbose)                       (expectedCombination.py:9)

Sums every `temp_exp_freq_<tag>_*.npy`, deletes ALL `temp_exp_freq_*.npy` (regardless of tag, as the
reference does, expectedCombination.py:38-39), normalises with float64 divide -> float32 on the device
(K4, expectedCombination.py:42) and saves `exp_freq_<tag>.npy`.
"""
from os import remove
from pathlib import Path
from sys import argv
from time import time

import numpy as np

from . import dist, helpers, session


def main(outputDirectory, storedExpInput, fileTag, verbose, backend=None):
    tTotal = time()
    outputDirPath, storedExpPath = Path(outputDirectory), Path(storedExpInput)
    if dist.rank() == 0:
        be = session.get_backend(backend)
        total = None
        for file in sorted(outputDirPath.glob("temp_exp_freq_{}_*.npy".format(fileTag))):
            part = be.to_device(np.load(file, allow_pickle=False))
            total = part if total is None else total + part          # integer adds: exact
        for file in outputDirPath.glob("temp_exp_freq_*.npy"):
            remove(file)
        if total is None:
            # the reference ends up normalising np.zeros((1, 1)) -> nan; keep the observable behaviour
            exp = (np.zeros((1, 1)) / np.sum(np.zeros((1, 1)))).astype(np.float32)
        else:
            exp = be.normalize(total.to(dtype=__import__("torch").int64)).cpu().numpy()
        np.save(storedExpPath, exp, allow_pickle=False)
        print("Total Time:", time() - tTotal) if verbose else print("    [Done]")
    dist.barrier()


if __name__ == "__main__":
    main(argv[1], argv[2], argv[3], helpers.strToBool(argv[4]))
