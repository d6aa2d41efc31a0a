class Patch:
    def __init__(self, *a, **k):
        pass
