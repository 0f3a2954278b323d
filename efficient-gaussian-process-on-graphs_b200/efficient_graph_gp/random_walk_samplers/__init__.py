from .sampler import RandomWalk, Graph

__all__ = ["RandomWalk", "Graph"]
