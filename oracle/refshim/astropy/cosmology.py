Planck18 = None
