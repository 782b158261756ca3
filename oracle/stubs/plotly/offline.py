def init_notebook_mode(*a, **k):
    pass


def iplot(*a, **k):
    pass
