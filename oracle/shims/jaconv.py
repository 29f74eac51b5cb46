jaconv = None
