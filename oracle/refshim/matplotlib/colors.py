def is_color_like(c):
    return True
