"""Stand-in: the reference imports `Grid` from tkinter and immediately shadows it."""


class Grid:
    pass
