class Figure:
    def __init__(self, *a, **k):
        raise RuntimeError("plotly stub: heatmap output is out of scope")


class Heatmap(Figure):
    pass
