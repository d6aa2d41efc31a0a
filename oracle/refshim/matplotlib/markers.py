class MarkerStyle:
    filled_markers = ('o', 'v', '^', '<', '>', '8', 's', 'p', '*', 'h', 'H', 'D', 'd', 'P', 'X')

    def __init__(self, *a, **k):
        pass
