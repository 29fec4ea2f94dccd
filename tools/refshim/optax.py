def adam(*a, **k):
    return None
