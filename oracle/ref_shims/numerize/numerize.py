def numerize(n, decimals=2):
    n = float(n)
    for div, suf in ((1e12, "T"), (1e9, "B"), (1e6, "M"), (1e3, "K")):
        if abs(n) >= div:
            return f"{round(n / div, decimals):g}{suf}"
    return f"{n:g}"
