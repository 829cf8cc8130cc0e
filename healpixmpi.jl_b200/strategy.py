"""Distribution strategies (reference src/strategy.jl:8,16): dispatch tags only."""


class Strategy:
    """Abstract distribution strategy (src/strategy.jl:8)."""


class RR(Strategy):
    """Round Robin: rank r owns m = r, r+P, ... and the ring pairs eq -+ (r + kP), the equator
    going to rank 0 (src/strategy.jl:16, src/alm.jl:112-119, src/map.jl:113-121)."""
