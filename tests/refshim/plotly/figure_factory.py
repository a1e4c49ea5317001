def create_annotated_heatmap(*a, **k):
    raise RuntimeError("plotly stub: heatmap output is out of scope (SURVEY.md section 2, row 8)")
