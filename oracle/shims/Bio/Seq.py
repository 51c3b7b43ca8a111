class Seq(str):
    """str subclass: `str(seq)`, `len(seq)`, slicing, `==` with str all behave as the scripts expect."""
    def __new__(cls, data=""):
        return str.__new__(cls, str(data))
