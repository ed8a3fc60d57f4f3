"""Import stub (pasero/preprocessing.py:20); identity normalisers, never reached by the ST path."""


def remove_non_printing_char(x):
    return x


def replace_unicode_punct(x):
    return x
