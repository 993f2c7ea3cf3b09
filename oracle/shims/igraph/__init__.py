"""Empty stand-in for ``igraph`` (absent here). The reference only does ``isinstance(X, ig.Graph)``
(`model.py:107`). Test infrastructure only."""


class Graph(object):
    pass
