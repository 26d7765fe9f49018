class NEProblem:
    pass
