"""B200-native Euler residual + RK4 (drop-in for PDESolver.jl's hot path)."""
__all__ = ["sbp", "mesh", "euler", "rk4", "lib"]
