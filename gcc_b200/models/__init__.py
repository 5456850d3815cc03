def __getattr__(name):
    if name == "GraphEncoder":
        from .graph_encoder import GraphEncoder
        return GraphEncoder
    raise AttributeError(name)


__all__ = ["GraphEncoder"]
