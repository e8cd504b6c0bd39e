def cdiv(a: int, b: int) -> int:
    return (a + b - 1) // b


KB = 1024
MB = 1024 * 1024
GB = 1024 * 1024 * 1024
TB = 1024 * 1024 * 1024 * 1024
