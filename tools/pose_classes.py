"""Developer tool: summarise a gpu_check json by pose class (horizontal poses 0-17, transition 18-35, looking down 36-59)."""
import json, sys
for f in sys.argv[1:]:
    r = json.load(open(f))
    def mean(lo, hi): 
        v = [x["p1_warm_ms"] for x in r if lo <= x["frame"] < hi]
        return sum(v) / max(1, len(v))
    tot = sum(x["p1_warm_ms"] + x["p2_warm_ms"] for x in r) / len(r)
    print(f"{f}: horiz {mean(0,18):.3f}  mid {mean(18,36):.3f}  down {mean(36,60):.3f}  | frame {tot:.3f} ms = {1000/tot:.0f} fps")
