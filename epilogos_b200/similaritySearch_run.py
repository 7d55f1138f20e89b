"""`simsearch` command line.  Mirror of similaritySearch_run.py of the reference (options :77-113, build :139-217, query
:234-287, block sizes :290-345).

    simsearch -b -s scores_x.txt.gz -o OUT [-w 25000] [-n 100] [-f -1] [--filter-score -1]
    simsearch -q chr1:1000000-1100000 -m OUT/simsearch.bed.gz -o QUERY_OUT

The reference can only build through SLURM: it submits similaritySearch_max_mean, nJobs x similaritySearch_calc and
similaritySearch_write as dependent batch jobs (:186-214).  Here the three stages run in this process, the distance
engine on the GPU; under `torchrun` the regions are split over the ranks exactly as the reference splits them over its
jobs (splitRows(regions, nJobs)[job], similaritySearch_calc.py:25) and rank 0 writes the result.  The SLURM-only
options (-j, -c, -p, -t, --mm-mem, --calc-mem, --write-mem) are accepted and ignored.
"""
import gzip
import re
from pathlib import Path
from time import time

import click
import numpy as np

from . import dist


def determineBinSize(scoresPath):
    """Bin width from the first row of the score file (similaritySearch_run.py:220-231)."""
    with gzip.open(scoresPath, "rt") if str(scoresPath).endswith(".gz") else open(scoresPath, "rt") as f:
        fields = f.readline().split("\t")
    return int(fields[2]) - int(fields[1])


_BLOCK_200 = {5000: 1, 10000: 2, 25000: 5, 50000: 10, 75000: 15, 100000: 20}
_BLOCK_20 = {500: 1, 1000: 2, 2500: 5, 5000: 10, 7500: 15, 10000: 20}


def determineBlockSize200(windowBP):
    """similaritySearch_run.py:319-345: every supported window reduces to 25 points."""
    if windowBP not in _BLOCK_200:
        raise ValueError("Error: window size must be either 5000, 10000, 25000, 50000, 75000, or 100000 (in bp)")
    return _BLOCK_200[windowBP]


def determineBlockSize20(windowBP):
    """similaritySearch_run.py:290-316."""
    if windowBP not in _BLOCK_20:
        raise ValueError("Error: window size must be either 500, 1000, 2500, 5000, 7500, or 10000 (in bp)")
    return _BLOCK_20[windowBP]


def generateRegionArr(query):
    """`chr:start-end` or a bed file of such regions -> object array [n, 3] (helpers.generateRegionArr, helpers.py:197-221)."""
    if re.fullmatch("chr[a-zA-z\\d]+:[\\d]+-[\\d]+", query):
        chrom, span = query.split(":")
        return np.array([[chrom, int(span.split("-")[0]), int(span.split("-")[1])]], dtype=object)
    if Path(query).is_file():
        rows = []
        with open(query) as f:
            for line in f:
                if line.strip():
                    c, s, e = line.rstrip("\n").split("\t")[:3]
                    rows.append([c, int(s), int(e)])
        return np.array(rows, dtype=object).reshape(len(rows), 3)
    raise ValueError("Please input valid query (region formatted as chr:start-end"
                     + "or path to bed file containing query regions)")


def buildSimSearch(scoresPath, outputDir, windowBP=-1, nDesiredMatches=100, filterState=-1, filterScore=-1.0):
    """similaritySearch_run.buildSimSearch (:139-217) without the batch system.  Returns the combined int32 index array
    on rank 0 (None elsewhere)."""
    from . import similaritySearch_calc, similaritySearch_max_mean, similaritySearch_write
    outputDir = Path(outputDir)
    binSize = determineBinSize(scoresPath)
    if binSize == 200:
        windowBP = 25000 if windowBP == -1 else windowBP
        windowBins, blockSize = int(windowBP / 200), determineBlockSize200(windowBP)
    elif binSize == 20:
        windowBP = 2500 if windowBP == -1 else windowBP
        windowBins, blockSize = int(windowBP / 20), determineBlockSize20(windowBP)
    else:
        raise ValueError("Similarity Search is only compatible with bins of size 200bp or 20bp")
    rank, world = dist.rank(), dist.world_size()

    if rank == 0:
        print("\n        STEP 1: Salient Region Selection", flush=True)
        similaritySearch_max_mean.main(outputDir, Path(scoresPath), windowBins, blockSize, windowBP, filterState, filterScore)
    dist.barrier()
    print("\n        STEP 2: Similarity Search Calculation", flush=True) if rank == 0 else None
    similaritySearch_calc.main(outputDir, windowBins, blockSize, 0, nDesiredMatches, world, rank)
    dist.barrier()
    if rank != 0:
        return None
    print("\n        STEP 3: Writing results", flush=True)
    similaritySearch_write.main(outputDir, windowBins, blockSize, world, nDesiredMatches)
    return np.load(outputDir / "simsearch_indices.npy", allow_pickle=True)


def querySimSearch(query, simSearchPath, outputDir):
    """similaritySearch_run.querySimSearch (:234-287): for every query range the FIRST region of the matches file that
    lies inside it; its matches go to `similarity_search_region_<chr>_<start>_<end>_recs.bed`.  Returns the files written."""
    print("\n\n\n        Reading in data...", flush=True); t = time()
    queryArr = generateRegionArr(query)
    regions = []
    with gzip.open(simSearchPath, "rt") as f:
        for line in f:
            c, s, e, recs = line.rstrip("\n").split("\t", 3)
            regions.append((c, int(s), int(e), recs))
    print("            Time:", format(time() - t, '.0f'), "seconds\n", flush=True)
    print("        Querying regions...", flush=True)
    written = []
    for chrom, start, end in queryArr:
        hit = next((r for r in regions if r[0] == chrom and r[1] >= start and r[2] <= end), None)
        if hit is None:
            print("            ValueError: Could not find region in given query range: {}:{}-{}\n".format(chrom, start, end))
            continue
        regionChr, regionStart, regionEnd, recs = hit
        outfile = Path(outputDir) / "similarity_search_region_{}_{}_{}_recs.bed".format(regionChr, regionStart, regionEnd)
        matches = recs[2:-2].split('", "')[1:]                      # brackets trimmed, the region itself dropped (:269-270)
        with open(outfile, "w+") as f:
            f.write("".join("{0[0]}\t{0[1]}\t{0[2]}\n".format(m.split(":")) for m in matches))
        written.append(outfile)
        print("            Found region {}:{}-{} within user query {}:{}-{}".format(regionChr, regionStart, regionEnd, chrom,
                                                                                    start, end))
        print("                See {} for matches\n".format(outfile), flush=True)
    return written


@click.command(context_settings=dict(help_option_names=["-h", "--help"]))
@click.option("-b", "--build", "buildBool", is_flag=True, help="Build the similarity search files needed to query regions")
@click.option("-s", "--scores", "scoresPath", type=str, help="Path to the scores file to be used in similarity search")
@click.option("-o", "--output-directory", "outputDir", required=True, type=str, help="Output directory")
@click.option("-w", "--window-bp", "windowBP", type=int, default=-1, help="Window size in bp [default: 25000]")
@click.option("-j", "--num-jobs", "nJobs", type=int, default=10, help="SLURM option of the reference (ignored)")
@click.option("-c", "--num-cores", "nCores", type=int, default=1, help="SLURM option of the reference (ignored)")
@click.option("-n", "--num-matches", "nDesiredMatches", type=int, default=100, show_default=True,
              help="Number of matches to be found for each query region")
@click.option("-f", "--filter-state", "filterState", type=int, default=-1,
              help="Regions whose max signal is in this state are removed; 0 = no filter [default: last state]")
@click.option("--filter-score", "filterScore", type=float, default=-1,
              help="Regions whose max signal is below this score are removed [default: -1 == no filtering]")
@click.option("-p", "--partition", "partition", type=str, help="SLURM option of the reference (ignored)")
@click.option("-t", "--tag", "jobTag", type=str, default="", help="SLURM option of the reference (ignored)")
@click.option("--mm-mem", "mmMem", type=str, default=10000, help="SLURM option of the reference (ignored)")
@click.option("--calc-mem", "calcMem", type=int, default=50000, help="SLURM option of the reference (ignored)")
@click.option("--write-mem", "writeMem", type=int, default=5000, help="SLURM option of the reference (ignored)")
@click.option("-q", "--query", "query", type=str, default="",
              help="Query region formatted as chr:start-end or path to a tab-separated bed file of query regions")
@click.option("-m", "--matches-file", "simSearchPath", type=str, help="Previously built simsearch.bed.gz to be queried")
def main(buildBool, scoresPath, outputDir, windowBP, nJobs, nCores, nDesiredMatches, filterState, filterScore, partition,
         jobTag, mmMem, calcMem, writeMem, query, simSearchPath):
    if not buildBool and query == "":
        raise ValueError("Either -b or -q flag must be used to run simsearch")
    elif buildBool and query != "":
        raise ValueError("Both -b and -q flags cannot be used at the same time")
    outputDir = Path(outputDir)
    if not outputDir.exists():
        outputDir.mkdir(parents=True, exist_ok=True)
    if not outputDir.is_dir():
        raise NotADirectoryError("Given path is not a directory: {}".format(str(outputDir)))
    if buildBool:
        dist.init_from_env()
        buildSimSearch(scoresPath, outputDir, windowBP, nDesiredMatches, filterState, filterScore)
    else:
        querySimSearch(query, simSearchPath, outputDir)


if __name__ == "__main__":
    main()
