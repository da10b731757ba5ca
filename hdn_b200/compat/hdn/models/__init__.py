"""Network blocks of the similarity / log-polar branches (mirror of hdn/models)."""
