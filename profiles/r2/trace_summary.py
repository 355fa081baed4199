"""Per-kernel windows of a partitioned world's tick from the -DGRIDDY_TRACE stamps (bench.py GRIDDY_DUMP_TRACE=dir).

Every stamp is %globaltimer of block 0 / thread 0 of a kernel right after its grid-dependency wait, i.e. the moment the
kernel before it on the stream has completed, so differences of consecutive stamps are whole kernel windows (launch gap,
waiting for a neighbour's flag and work).  usage: python profiles/r2/trace_summary.py dir world
"""
import sys
import numpy as np

A, S, U, L, HP, HU, EP, EF, IU, IG = range(10)   # records.cuh: TR_ADVANCE, TR_SLOW, TR_UNLINK, TR_LINK, TR_HALO_PACK, ...


def main(d, world):
    for r in range(world):
        t = np.load(f"{d}/trace_n{world}_rank{r}.npy").astype(np.int64)
        ticks = np.argsort(t[:, A])          # rows are tick mod 64: order them by time
        t = t[ticks]
        t = t[t[:, A] > 0][-45:]
        nxt = t[1:, A]
        t = t[:-1]
        if (t[:, HU] > 0).all():   # halo and immigrants taken in by kernels of their own
            w = {"k_advance (incl. halo pack)": t[:, HU] - t[:, A], "k_halo_unpack (wait + scatter)": t[:, S] - t[:, HU],
                 "k_slow (incl. publish)": t[:, U] - t[:, S], "k_unlink": t[:, IU] - t[:, U],
                 "k_immig_unpack": t[:, L] - t[:, IU], "  of which waiting for flags": t[:, IG] - t[:, IU]}
        else:                      # ... by the last blocks of k_advance / k_unlink
            w = {"k_advance (halo out and in)": t[:, S] - t[:, A], "k_slow (incl. publish)": t[:, U] - t[:, S],
                 "k_unlink (and immigrants)": t[:, L] - t[:, U], "  migrant flags seen after k_unlink's start": t[:, IG] - t[:, U]}
        w.update({"k_link": nxt - t[:, L], "tick": nxt - t[:, A], "emigrant flags raised after k_slow's start": t[:, EF] - t[:, S]})
        print(f"rank {r}: " + "; ".join(f"{k} {np.median(v) / 1e3:.1f} us" for k, v in w.items()))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
