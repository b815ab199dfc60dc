"""bench.py --workload ring|mat: BASELINE configs[3] (sample-shard ring) and configs[4] (.mat count matrices)."""


def main(args, rank, world, local_rank, emit, log, ClockSampler, measured_peaks):
    raise SystemExit(f"bench.py: --workload {args.workload} is not available in this build")
