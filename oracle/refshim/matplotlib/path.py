class Path:
    MOVETO = LINETO = CLOSEPOLY = 0

    def __init__(self, *a, **k):
        pass
