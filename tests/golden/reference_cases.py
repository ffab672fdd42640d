"""Networks of the reference-generated golden fixtures (tests/golden/reference_steps.npz).

A case is a list of layer tuples that three builders understand: the reference's own classes through
oracle/ref.py (make_reference_fixtures.py, which writes the fixture), the numpy oracle (CPU check of the
fixture) and the product's Python API (the `-m gpu` test).  Layer tuples:
    ("hyperplane", input, output, wname, bname)      ("actf", kind[, {params}])
    ("rewrap", dims)   ("convolution", kernel, n, wname[, step])   ("convolution_bias", n, wname)
    ("max_pooling", kernel)   ("flatten",)   ("prelu", size, wname, scalar)
"""

CASES = {
    # name: (layers, input size, output size, bunch, loss, target kind)
    "mlp_tanh_logistic_mcce": ([("hyperplane", 20, 16, "w1", "b1"), ("actf", "tanh"),
                                ("hyperplane", 16, 12, "w2", "b2"), ("actf", "logistic"),
                                ("hyperplane", 12, 5, "w3", "b3"), ("actf", "log_softmax")],
                               20, 5, 9, "multi_class_cross_entropy", "onehot"),
    "mlp_relu_mcce": ([("hyperplane", 20, 24, "w1", "b1"), ("actf", "relu"),
                       ("hyperplane", 24, 24, "w2", "b2"), ("actf", "relu"),
                       ("hyperplane", 24, 5, "w3", "b3"), ("actf", "log_softmax")],
                      20, 5, 9, "multi_class_cross_entropy", "onehot"),
    "mlp_softmax_mse": ([("hyperplane", 20, 16, "w1", "b1"), ("actf", "logistic"),
                         ("hyperplane", 16, 5, "w2", "b2"), ("actf", "softmax")],
                        20, 5, 7, "mse", "dense"),
    "mlp_extra_actfs_mse": ([("hyperplane", 18, 14, "w1", "b1"), ("actf", "softplus"),
                             ("hyperplane", 14, 14, "w2", "b2"), ("actf", "softsign"),
                             ("hyperplane", 14, 12, "w3", "b3"), ("actf", "leaky_relu", {"leak": 0.2}),
                             ("hyperplane", 12, 10, "w4", "b4"), ("actf", "hardtanh", {"inf": -0.5, "sup": 0.5}),
                             ("prelu", 10, "a5", False),
                             ("hyperplane", 10, 4, "w6", "b6"), ("actf", "linear")],
                            18, 4, 8, "mse", "dense"),
    "mlp_log_logistic_ce": ([("hyperplane", 12, 9, "w1", "b1"), ("actf", "tanh"),
                             ("hyperplane", 9, 6, "w2", "b2"), ("actf", "log_logistic")],
                            12, 6, 10, "cross_entropy", "binary"),
    "digits_mlp_mcce": ([("hyperplane", 256, 256, "w1", "b1"), ("actf", "tanh"),
                         ("hyperplane", 256, 128, "w2", "b2"), ("actf", "tanh"),
                         ("hyperplane", 128, 10, "w3", "b3"), ("actf", "log_softmax")],
                        256, 10, 32, "multi_class_cross_entropy", "onehot"),
    "conv_pool_mcce": ([("rewrap", (1, 16, 16)),
                        ("convolution", (1, 5, 5), 4, "cw1"), ("convolution_bias", 4, "cb1"), ("actf", "relu"),
                        ("max_pooling", (1, 2, 2)),
                        ("convolution", (4, 3, 3), 6, "cw2"), ("convolution_bias", 6, "cb2"), ("actf", "tanh"),
                        ("max_pooling", (1, 2, 2)),
                        ("flatten",),
                        ("hyperplane", 24, 10, "w3", "b3"), ("actf", "log_softmax")],
                       256, 10, 6, "multi_class_cross_entropy", "onehot"),
    "conv_stride2_mse": ([("rewrap", (2, 9, 9)),
                          ("convolution", (2, 3, 3), 5, "cw1", (1, 2, 2)), ("convolution_bias", 5, "cb1"),
                          ("actf", "tanh"), ("flatten",),
                          ("hyperplane", 80, 3, "w2", "b2"), ("actf", "linear")],
                         162, 3, 5, "mse", "dense"),
}


def weight_names(layers):
    names = []
    for l in layers:
        if l[0] == "hyperplane":
            names += [l[3], l[4]]
        elif l[0] in ("convolution",):
            names.append(l[3])
        elif l[0] == "convolution_bias":
            names.append(l[2])
        elif l[0] == "prelu":
            names.append(l[2])
    return names


def build_reference(R, layers, input_size, output_size):
    s = R.stack()
    for l in layers:
        k = l[0]
        if k == "hyperplane":
            R.push(s, R.hyperplane(l[1], l[2], l[3], l[4]))
        elif k == "actf":
            p = l[2] if len(l) > 2 else {}
            R.push(s, R.actf(l[1], p.get("leak", p.get("inf", 0.0)), p.get("sup", 0.0)))
        elif k == "rewrap":
            R.push(s, R.rewrap(list(l[1])))
        elif k == "convolution":
            R.push(s, R.convolution(list(l[1]), l[2], l[3], list(l[4]) if len(l) > 4 else None))
        elif k == "convolution_bias":
            R.push(s, R.convolution_bias(3, l[1], l[2]))
        elif k == "max_pooling":
            R.push(s, R.max_pooling(list(l[1])))
        elif k == "flatten":
            R.push(s, R.flatten())
        elif k == "prelu":
            R.push(s, R.prelu(l[1], l[2], l[3]))
        else:
            raise ValueError(k)
    return R.Net(s, input_size, output_size)


def build_oracle(A, layers, input_size, weights):
    o = A.Stack()
    for l in layers:
        k = l[0]
        if k == "hyperplane":
            A.hyperplane(o, l[1], l[2], l[3], l[4])
        elif k == "actf":
            o.push(A.Actf(l[1], **(l[2] if len(l) > 2 else {})))
        elif k == "rewrap":
            o.push(A.Rewrap(tuple(l[1])))
        elif k == "convolution":
            o.push(A.Convolution(tuple(l[1]), l[2], l[3], tuple(l[4]) if len(l) > 4 else None))
        elif k == "convolution_bias":
            o.push(A.ConvolutionBias(l[1], l[2]))
        elif k == "max_pooling":
            o.push(A.MaxPooling(tuple(l[1])))
        elif k == "flatten":
            o.push(A.Flatten())
        elif k == "prelu":
            o.push(A.PReLU(l[1], l[2], l[3]))
        else:
            raise ValueError(k)
    o.build(input_size, weights)
    return o


def build_product(ann, layers):
    c = ann.components
    net = c.stack(name="stack")
    for i, l in enumerate(layers):
        k, nm = l[0], "c%d" % i
        if k == "hyperplane":
            net.push(c.hyperplane(input=l[1], output=l[2], name=nm, dot_product_name=nm + "w", bias_name=nm + "b",
                                  dot_product_weights=l[3], bias_weights=l[4]))
        elif k == "actf":
            p = l[2] if len(l) > 2 else {}
            net.push(getattr(c.actf, l[1])(name=nm, **p))
        elif k == "rewrap":
            net.push(c.rewrap(size=tuple(l[1]), name=nm))
        elif k == "convolution":
            net.push(c.convolution(kernel=tuple(l[1]), n=l[2], name=nm, weights=l[3],
                                   step=tuple(l[4]) if len(l) > 4 else None))
        elif k == "convolution_bias":
            net.push(c.convolution_bias(n=l[1], ndims=3, name=nm, weights=l[2]))
        elif k == "max_pooling":
            net.push(c.max_pooling(kernel=tuple(l[1]), name=nm))
        elif k == "flatten":
            net.push(c.flatten(name=nm))
        elif k == "prelu":
            net.push(c.actf.prelu(size=l[1], scalar=l[3], name=nm, weights=l[2]))
        else:
            raise ValueError(k)
    return net
