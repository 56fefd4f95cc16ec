from . import numerize  # noqa: F401  (the reference does `from numerize import numerize; numerize.numerize(x)`)
