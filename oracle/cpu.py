"""Host-core accounting for the CPU legs (TEST INFRASTRUCTURE).

``os.cpu_count()`` reports the machine (128 on the GPU box) but the container is limited by a cgroup
CPU quota (16 CPUs there); running eager torch with 128 threads on a 16-CPU quota is ~8x slower than
with 16.  ``usable_cores()`` = min(scheduler affinity, cgroup quota)."""
import math
import os


def usable_cores() -> int:
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max",):
        try:
            quota, period = open(path).read().split()[:2]
            if quota != "max":
                n = min(n, max(1, math.ceil(int(quota) / int(period))))
        except Exception:
            pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            n = min(n, max(1, math.ceil(q / p)))
    except Exception:
        pass
    return max(1, n)
