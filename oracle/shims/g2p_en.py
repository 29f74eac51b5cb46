# empty stub
